// TEMPORARY link stubs while subsystems are being brought up.
#include "engine.h"
namespace dv {
int mix_init(Engine*) { return DV_OK; }
void mix_free(Engine*) {}
int lg_init(Engine*) { return DV_OK; }
void lg_free(Engine*) {}
int store_init(Engine*) { return DV_OK; }
void store_free(Engine*) {}
void comm_free(Engine*) {}
}
