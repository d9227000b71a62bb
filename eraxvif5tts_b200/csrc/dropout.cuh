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

// Site 2, the dropout inside F.scaled_dot_product_attention (model/modules.py:490), has its own, cheaper stream: the attention kernels
// spend ~5 instructions per score, so a 64-bit splitmix per 4 scores tripled their cost.  One Philox-2x32 block (7 rounds of
// mul.wide + 3-input xor) serves the 8 consecutive keys of one query row: group index g = ((b*H + h)*n + query) * ceil(n/8) + key/8,
// counter = (lo32(g), hi32(g) ^ chi), key = `key`; element `key & 7` owns byte (key & 3) of word (key & 7) >> 2 and is KEPT iff
// (byte & 0x7f) >= t7, t7 = round(p * 128): the probability is quantised to 1/128 (flash-attention quantises to 1/256) and the kept
// values are scaled by 128 / (128 - t7), so the expectation is exact.  attn_drop_words returns the two words with the keep flag in
// the top bit of every byte (a carry-free packed add), ready for a sign-replicating PRMT.
struct AttnDrop {
  uint32_t addc;     // (128 - t7) in every byte; 0 = dropout off
  float scale;       // 128 / (128 - t7)
  float log2_scale;
  uint32_t key, chi;
};
__device__ __forceinline__ void attn_drop_words(const AttnDrop& d, uint64_t g, uint32_t& w0, uint32_t& w1) {
  uint32_t c0 = (uint32_t)g, c1 = (uint32_t)(g >> 32) ^ d.chi, k = d.key;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint64_t pr = (uint64_t)c0 * 0xD256D193u;
    c0 = (uint32_t)(pr >> 32) ^ k ^ c1;
    c1 = (uint32_t)pr;
    k += 0x9E3779B9u;
  }
  w0 = (c0 & 0x7f7f7f7fu) + d.addc;
  w1 = (c1 & 0x7f7f7f7fu) + d.addc;
}
// 32-bit AND mask of a packed bf16 pair whose elements own bytes (2*pair, 2*pair + 1) of `w`: the byte's top bit replicated
__device__ __forceinline__ uint32_t attn_drop_pair_mask(uint32_t w, int pair) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(m) : "r"(w), "r"(pair ? 0xBBAAu : 0x9988u));
  return m;
}

// the four keep flags of one word of attn_drop_words (top bit of every byte) as a nibble: bit i = byte i
__device__ __forceinline__ uint32_t attn_drop_nibble(uint32_t w) {
  return (((w >> 7) & 0x01010101u) * 0x01020408u) >> 24;  // bit 8i lands on bit 24 + i; the partial products never collide
}
// 32 x 32 bit-matrix transpose across a warp: lane q holds row q (bit k = element (q, k)) -> lane k holds column k (bit q)
__device__ __forceinline__ uint32_t warp_bit_transpose(uint32_t x, int lane) {
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    // m: the columns whose index has bit j clear.  The upper row of a pair (lane bit j clear) keeps those and takes the partner's,
    // shifted up by j; the lower row keeps the others and takes the partner's others, shifted down: the off-diagonal j x j blocks swap
    const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t q = __shfl_xor_sync(0xffffffffu, x, j);
    x = (lane & j) ? ((x & ~m) | ((q & ~m) >> j)) : ((x & m) | ((q & m) << j));
  }
  return x;
}

AttnDrop make_attn_drop(float p, uint64_t seed, int layer);  // host (train_kernels.cu): the stream of one DiT block; p <= 0: off
AttnDrop attn_drop_for_layer(int layer);  // host: f5b_train_set_attn_dropout's stream (training drivers)

}  // namespace f5b
