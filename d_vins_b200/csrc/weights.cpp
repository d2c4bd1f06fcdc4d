// DVWGT001 weight-file reader (format defined in oracle/weights.py; written by any converter of real checkpoints).
#include <stdio.h>
#include <string.h>

#include "engine.h"

namespace dv {

int load_weight_file(const char* path, WeightMap* out) {
  FILE* f = fopen(path, "rb");
  if (!f) { set_error(std::string("cannot open weight file: ") + path); return DV_ERR_WEIGHTS; }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf((size_t)sz);
  if (fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); set_error("short read on weight file"); return DV_ERR_WEIGHTS; }
  fclose(f);
  if (sz < 12 || memcmp(buf.data(), "DVWGT001", 8) != 0) { set_error("bad weight file magic"); return DV_ERR_WEIGHTS; }
  size_t p = 8;
  auto rd32 = [&](uint32_t* v) { if (p + 4 > (size_t)sz) return false; memcpy(v, &buf[p], 4); p += 4; return true; };
  auto rd64 = [&](uint64_t* v) { if (p + 8 > (size_t)sz) return false; memcpy(v, &buf[p], 8); p += 8; return true; };
  uint32_t n = 0;
  if (!rd32(&n)) { set_error("truncated weight header"); return DV_ERR_WEIGHTS; }
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t ln = 0, nd = 0;
    if (!rd32(&ln) || p + ln > (size_t)sz) { set_error("truncated weight header"); return DV_ERR_WEIGHTS; }
    std::string name(reinterpret_cast<const char*>(&buf[p]), ln);
    p += ln;
    if (!rd32(&nd) || nd > 8) { set_error("bad ndim in weight header"); return DV_ERR_WEIGHTS; }
    HostTensor t;
    for (uint32_t d = 0; d < nd; ++d) {
      uint32_t v;
      if (!rd32(&v)) { set_error("truncated dims"); return DV_ERR_WEIGHTS; }
      t.dims.push_back((int)v);
    }
    uint64_t off = 0, nb = 0;
    if (!rd64(&off) || !rd64(&nb) || off > (uint64_t)sz || nb > (uint64_t)sz - off || nb != (uint64_t)t.numel() * 4) {
      set_error("bad tensor extent in weight file: " + name);
      return DV_ERR_WEIGHTS;
    }
    t.data.resize((size_t)t.numel());
    memcpy(t.data.data(), &buf[off], nb);
    (*out)[name] = std::move(t);
  }
  return DV_OK;
}

}  // namespace dv
