/*
 * TEST INFRASTRUCTURE -- CPU oracle for the CFG merge + sampling tail of the decode step.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this; the product
 * path (ospo_b200/) never does.
 *
 * Restates ospo/wrapper/image_generation.py:157-163 (== ospo/inference.py:148-154):
 *     logit_cond = logits[0::2]; logit_uncond = logits[1::2]                      :157-158
 *     logits = logit_uncond + cfg_weight * (logit_cond - logit_uncond)            :160
 *     probs  = softmax(logits / temperature)                                      :161
 *     next_token = multinomial(probs, 1)                                          :163
 * with two documented specialisations that make the result reproducible bit for bit:
 *   (1) merge_mode 0 rounds to bf16 after every elementwise op, which is what the bf16 tensors of
 *       the reference do (merge_mode 1 keeps fp32);
 *   (2) torch.multinomial draws from the global Philox stream and cannot be reproduced by another
 *       kernel, so sampling is defined as inverse-CDF on caller-supplied uniforms:
 *           id = min{ k : cdf_k > u * Z },  weights w_v = e^(t_v) / 2^K  (K = max_v rint(t_v log2 e)), Z = sum w,
 *       with a fully specified fp32 exp and a fixed summation order (tiles of 128 codes, segments of 32
 *       codes summed as a stride-16/8/4/2/1 butterfly tree, 16 groups x 32 segments and the 16 groups summed sequentially).
 * The probabilities w/Z are checked against the real reference's `probs` in tests (golden
 * tests/golden/cfg_ref.npz); the CUDA kernel is checked against this file bit for bit.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC cfg_sample.c -lm
 * (fp contraction must stay off so every line below is one IEEE-754 operation).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define SEG 32
#define GRP 32

static float bf16_to_f32(uint16_t b) {
  uint32_t u = ((uint32_t)b) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static float bf16_round(float f) { /* round-to-nearest-even to bf16, returned as float */
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return f; /* NaN */
  u += 0x7FFFu + ((u >> 16) & 1u);
  u &= 0xFFFF0000u;
  memcpy(&f, &u, 4);
  return f;
}

/* ---- softmax weights with a power-of-two reference --------------------------------------------------
 * For a merged logit t:  y = t * log2(e);  n = rint(y);  r = t - n ln2 (Cody-Waite, two fma);
 * e^t = P(r) * 2^n with P a degree-6 polynomial.  Weights are kept RELATIVE to an integer exponent K:
 *     w = P(r) * 2^(n - K)         (0 if n - K < -120)
 * so that changing K is an exact power-of-two rescale.  That lets a GPU kernel compute the weights of a
 * 128-code tile against the tile's own K_tile (while the tile is still in tensor memory) and rescale the
 * segment sums later, bit for bit the same as this file.  Every line is one IEEE-754 fp32 operation. */
static float exp_parts(float t, float* n_out) { /* returns P(r), writes n */
  float y = t * 1.4426950408889634f;
  if (!(y >= -1.0e4f)) y = -1.0e4f;
  if (!(y <= 1.0e4f)) y = 1.0e4f;
  float n = rintf(y);
  float r = fmaf(n, -0.693145751953125f, t);
  r = fmaf(n, -1.42860682030941723212e-6f, r);
  float p = 1.3888888888888889e-03f;
  p = fmaf(p, r, 8.3333333333333332e-03f);
  p = fmaf(p, r, 4.1666666666666664e-02f);
  p = fmaf(p, r, 1.6666666666666666e-01f);
  p = fmaf(p, r, 0.5f);
  p = fmaf(p, r, 1.0f);
  p = fmaf(p, r, 1.0f);
  *n_out = n;
  return p;
}

static float pow2_factor(float e) { /* 2^e for integer-valued e <= 0; 0 below -120 */
  if (e < -120.0f) return 0.0f;
  uint32_t sb = (uint32_t)((int)e + 127) << 23;
  float f;
  memcpy(&f, &sb, 4);
  return f;
}

float ospo_oracle_exp_det(float x) { /* e^x for x <= 0 (kept for tests of the polynomial) */
  float n;
  float p = exp_parts(x, &n);
  return p * pow2_factor(n);
}

static float merge_one(float lc, float lu, float w, float T, int merge_mode) {
  if (merge_mode == 0) {
    float d = bf16_round(lc - lu);
    float e = bf16_round(w * d);
    float m = bf16_round(lu + e);
    return bf16_round(m / T);
  } else {
    float d = lc - lu;
    float e = w * d;
    float m = lu + e;
    return m / T;
  }
}

/* logits: bf16 bit patterns [2P, V]; merged: fp32 [P, V] */
void ospo_oracle_cfg_merge(const uint16_t* logits, int P, int V, float w, float T, int merge_mode, float* merged) {
  for (int p = 0; p < P; ++p) {
    const uint16_t* lc = logits + (size_t)(2 * p) * V;
    const uint16_t* lu = lc + V;
    for (int v = 0; v < V; ++v) merged[(size_t)p * V + v] = merge_one(bf16_to_f32(lc[v]), bf16_to_f32(lu[v]), w, T, merge_mode);
  }
}

/* merged: fp32 [P, V] (V multiple of TILE and of SEG*GRP); uniforms [P]; ids [P];
 * weights_out (optional) [P, V] weights relative to the global exponent K; z_out (optional) [P].
 * Order of operations (mirrored exactly by the CUDA kernels):
 *   tile (128 codes):  K_tile = max n;  u_v = P(r_v) * 2^(n_v - K_tile)
 *   segment (32 codes): S = butterfly tree sum of u: x[j] += x[j+16] (j<16), then strides 8, 4, 2, 1
 *   K = max K_tile;  S' = S * 2^(K_tile - K)
 *   group (32 segments): sequential sum of S';  Z = sequential sum of the 16 group sums
 *   descent group -> segment -> code, `base` carried along; in-segment weights are u * 2^(K_tile - K) */
#define TILE 128
int ospo_oracle_sample_merged(const float* merged, int P, int V, const float* uniforms, int greedy, int64_t* ids,
                              float* weights_out, float* z_out) {
  if (V % (SEG * GRP) != 0 || V % TILE != 0 || V / SEG > 4096 || V > (1 << 20)) return -1;
  const int nseg = V / SEG, ngrp = nseg / GRP, ntile = V / TILE;
  static float ubuf[1 << 20], pbuf[1 << 20], nbuf[1 << 20];
  static float seg_sum[4096], grp_sum[128], tile_k[8192];
  for (int p = 0; p < P; ++p) {
    const float* t = merged + (size_t)p * V;
    if (greedy) {
      float gmax = t[0];
      int garg = 0;
      for (int v = 1; v < V; ++v) {
        if (t[v] > gmax) { gmax = t[v]; garg = v; }
      }
      ids[p] = garg;
      continue;
    }
    for (int v = 0; v < V; ++v) pbuf[v] = exp_parts(t[v], &nbuf[v]);
    float K = -INFINITY;
    for (int tl = 0; tl < ntile; ++tl) {
      float kt = nbuf[tl * TILE];
      for (int j = 1; j < TILE; ++j) kt = nbuf[tl * TILE + j] > kt ? nbuf[tl * TILE + j] : kt;
      tile_k[tl] = kt;
      if (kt > K) K = kt;
      for (int j = 0; j < TILE; ++j) ubuf[tl * TILE + j] = pbuf[tl * TILE + j] * pow2_factor(nbuf[tl * TILE + j] - kt);
    }
    for (int s = 0; s < nseg; ++s) {
      float x[SEG];
      for (int j = 0; j < SEG; ++j) x[j] = ubuf[s * SEG + j];
      for (int h = SEG / 2; h >= 1; h >>= 1) /* butterfly order: strides 16, 8, 4, 2, 1 */
        for (int j = 0; j < h; ++j) x[j] = x[j] + x[j + h];
      seg_sum[s] = x[0] * pow2_factor(tile_k[s / (TILE / SEG)] - K);
    }
    for (int g = 0; g < ngrp; ++g) {
      float acc = 0.0f;
      for (int j = 0; j < GRP; ++j) acc = acc + seg_sum[g * GRP + j];
      grp_sum[g] = acc;
    }
    float Z = 0.0f;
    for (int g = 0; g < ngrp; ++g) Z = Z + grp_sum[g];
    if (weights_out)
      for (int v = 0; v < V; ++v) weights_out[(size_t)p * V + v] = ubuf[v] * pow2_factor(tile_k[v / TILE] - K);
    if (z_out) z_out[p] = Z;
    const float target = uniforms[p] * Z;
    float base = 0.0f;
    int g = 0;
    for (; g < ngrp - 1; ++g) {
      float nxt = base + grp_sum[g];
      if (nxt > target) break;
      base = nxt;
    }
    int sg = 0;
    for (; sg < GRP - 1; ++sg) {
      float nxt = base + seg_sum[g * GRP + sg];
      if (nxt > target) break;
      base = nxt;
    }
    const int segi = g * GRP + sg;
    const float f = pow2_factor(tile_k[segi / (TILE / SEG)] - K);
    int j = 0;
    for (; j < SEG - 1; ++j) {
      float nxt = base + ubuf[segi * SEG + j] * f;
      if (nxt > target) break;
      base = nxt;
    }
    ids[p] = (int64_t)segi * SEG + j;
  }
  return 0;
}

int ospo_oracle_cfg_sample(const uint16_t* logits, int P, int V, float w, float T, int merge_mode, const float* uniforms,
                           int greedy, int64_t* ids, float* merged_out, float* weights_out, float* z_out) {
  static float mbuf[64 * 16384];
  if ((size_t)P * V > sizeof(mbuf) / sizeof(float)) return -2;
  float* merged = merged_out ? merged_out : mbuf;
  ospo_oracle_cfg_merge(logits, P, V, w, T, merge_mode, merged);
  return ospo_oracle_sample_merged(merged, P, V, uniforms, greedy, ids, weights_out, z_out);
}
