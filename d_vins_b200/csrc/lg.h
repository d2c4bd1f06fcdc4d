// LightGlue internal interface (shared by lg.cu and store.cu).
#pragma once
#include <stdint.h>

namespace dv {

struct Engine;

// One image's tokens: device pointers to pixel keypoints [n,2] and descriptors [n,256] (fp32).
struct LgSeg {
  const float* kpts;
  const float* desc;
  int n;        // keypoints
  int w, h;     // image size used for the caller-side normalisation (deep_net.cpp:839-841)
  int off;      // first row in the packed token buffers (filled by lg_run)
};

int lg_run(Engine* e, int P, const LgSeg* segs);   // segs [2P]: (query, old) per pair; results stay on device
int lg_fetch_batch(Engine* e, int P, int cap, const int* slot, int32_t* matches, float* mscores, int32_t* k_out);
int lg_fetch(Engine* e, int p, int cap, int32_t* matches, float* mscores, float* mk0, float* mk1, int32_t* k_out);

}  // namespace dv
