// hash_host.h -- host-side pieces of videocompare's hashes shared by b200vfx.cu and hash_host.cpp (not part of the ABI)
#pragma once
#include <vector>

namespace b200vfx {

// image::imageops::sample (image 0.25.10, recalled): normalised Lanczos3 tap weights of every output sample when
// `in_len` source samples are resized to `out_len`.  taps: out_len rows of max_taps floats; meta[o] = {left, n}.
struct ResizeTaps {
  int max_taps = 0;
  std::vector<float> taps;
  std::vector<int> left, count;
};
ResizeTaps make_resize_taps(int in_len, int out_len);

}  // namespace b200vfx
