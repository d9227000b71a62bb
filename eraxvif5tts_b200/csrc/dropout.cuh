// Counter-based dropout masks shared by the training sweeps (train_kernels.cu) and the attention kernels.
#pragma once
#include <stdint.h>

namespace f5b {

// Dropout (train mode of the reference's DiT: FeedForward's Dropout after GELU, model/modules.py:342-353, and the Dropout behind
// attention's to_out, :436-440).  Counter-based: the keep decision of element idx is a pure function of (key, idx), so the backward
// regenerates the forward's mask instead of storing it.  One splitmix64 hash serves 4 consecutive elements (16 bits each).
// thr16 == 0 switches it off.  Site 2 is the dropout inside F.scaled_dot_product_attention (:490): attention.cu / attention_bwd.cu.
struct Drop {
  uint32_t thr16;  // drop if lane bits < thr16 (= p * 65536)
  float scale;     // 1 / (1 - p)
  uint64_t key;
};
__device__ __forceinline__ uint64_t drop_hash(uint64_t key, uint64_t idx4) {
  uint64_t z = idx4 * 0x9E3779B97F4A7C15ull + key;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// multipliers (0 or 1/(1-p)) of the 4 consecutive elements starting at element index idx (a multiple of 4)
__device__ __forceinline__ void drop_mult4(const Drop& d, uint64_t idx, float (&m)[4]) {
  const uint64_t z = drop_hash(d.key, idx >> 2);
#pragma unroll
  for (int k = 0; k < 4; ++k) m[k] = ((uint32_t)(z >> (16 * k)) & 0xFFFFu) >= d.thr16 ? d.scale : 0.f;
}


// multiplier (0 or 1/(1-p)) of ONE element: `z` = drop_hash of the element's group of 4, `lane` = element index & 3
__device__ __forceinline__ float drop_mult1(const Drop& d, uint64_t z, int lane) {
  return ((uint32_t)(z >> (16 * lane)) & 0xFFFFu) >= d.thr16 ? d.scale : 0.f;
}

Drop drop_for_site(int layer, int site);  // host (train_kernels.cu): the mask stream of (layer, site) under f5b_train_set_*dropout

}  // namespace f5b
