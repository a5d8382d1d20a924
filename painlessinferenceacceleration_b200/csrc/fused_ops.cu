// Elementwise pieces of the LOOKAHEAD verify forward (sm_100a), all HBM-bound, 16-byte vectorised:
//   k_rmsnorm           models/llama/modeling_llama.py:76-90   (+ the residual add of the decoder layer :340-352)
//   k_rope_kv_append    :156-169 apply_rotary_pos_emb at the tree positions of :587, and the KV-cache append
//                       that replaces the reference's per-step torch.cat (:265-268)
//   k_silu_mul          :185-186
//   k_embed_gather      :582
// Rounding points follow the reference's bf16 eager arithmetic (each torch op rounds to bf16) so that
// the verify logits stay as close to the reference's as a different GEMM order allows.
#include <cuda_bf16.h>

#include "common.cuh"

namespace pia {
namespace fused {

__device__ __forceinline__ float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

union Pack8 { uint4 u; __nv_bfloat16 h[8]; };

// one CTA per row; hidden % 8 == 0
// x comes either as bf16 [rows, hidden] or as `n_parts` fp32 split-K slices [n_parts][part_rows][hidden] of the
// producing GEMM (summed in slice order and rounded to bf16 first, i.e. what a bf16 GEMM output would hold)
__device__ __forceinline__ Pack8 load_x8(const __nv_bfloat16 *x, const float *parts, int n_parts, long long part_stride,
                                         long long off) {
  Pack8 a;
  if (parts == nullptr) { a.u = *reinterpret_cast<const uint4 *>(x + off); return a; }
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < n_parts; ++s) {
    const float4 lo = *reinterpret_cast<const float4 *>(parts + s * part_stride + off);
    const float4 hi = *reinterpret_cast<const float4 *>(parts + s * part_stride + off + 4);
    acc[0] += lo.x; acc[1] += lo.y; acc[2] += lo.z; acc[3] += lo.w;
    acc[4] += hi.x; acc[5] += hi.y; acc[6] += hi.z; acc[7] += hi.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) a.h[j] = __float2bfloat16_rn(acc[j]);
  return a;
}

__global__ void __launch_bounds__(512) k_rmsnorm(const __nv_bfloat16 *x, const float *parts, int n_parts,
                                                 long long part_stride, const __nv_bfloat16 *res_in,
                                                 const __nv_bfloat16 *w, float eps, int hidden,
                                                 __nv_bfloat16 *res_out, __nv_bfloat16 *y) {
  __shared__ float red[16];
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x, tid = threadIdx.x;
  const int nvec = hidden >> 3;
  const long long row_off = (long long)row * hidden;
  const uint4 *rv = res_in ? reinterpret_cast<const uint4 *>(res_in + (long long)row * hidden) : nullptr;
  uint4 *rov = res_out ? reinterpret_cast<uint4 *>(res_out + (long long)row * hidden) : nullptr;
  float ss = 0.f;
  // hidden <= 8 * 512 * 2 : keep up to two vectors per thread in registers
  Pack8 keep[2];
  int cnt = 0;
  for (int v = tid; v < nvec; v += 512) {
    Pack8 a = load_x8(x, parts, n_parts, part_stride, row_off + v * 8);
    if (rv) {
      Pack8 r; r.u = rv[v];
#pragma unroll
      for (int j = 0; j < 8; ++j) a.h[j] = __float2bfloat16_rn(__bfloat162float(a.h[j]) + __bfloat162float(r.h[j]));
    }
    if (rov) rov[v] = a.u;
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float f = __bfloat162float(a.h[j]); ss += f * f; }
    if (cnt < 2) keep[cnt] = a;
    ++cnt;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL, ss, o);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) tot += red[i];
  const float inv = rsqrtf(tot / (float)hidden + eps);
  const uint4 *wv = reinterpret_cast<const uint4 *>(w);
  uint4 *yv = reinterpret_cast<uint4 *>(y + (long long)row * hidden);
  cnt = 0;
  for (int v = tid; v < nvec; v += 512) {
    Pack8 a;
    if (cnt < 2) a = keep[cnt];
    else {  // (only for hidden > 8192) recompute the residual sum
      a = load_x8(x, parts, n_parts, part_stride, row_off + v * 8);
      if (rv) { Pack8 r; r.u = rv[v];
#pragma unroll
        for (int j = 0; j < 8; ++j) a.h[j] = __float2bfloat16_rn(__bfloat162float(a.h[j]) + __bfloat162float(r.h[j])); }
    }
    ++cnt;
    Pack8 ww; ww.u = wv[v];
    Pack8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.h[j] = __float2bfloat16_rn(__bfloat162float(ww.h[j]) * (__bfloat162float(a.h[j]) * inv));
    yv[v] = o.u;
  }
}

// grid = batch * rows_per_slot rows (pia_slots_t); thread = one (head, 8-wide d chunk) of q | k | v
__global__ void __launch_bounds__(256) k_rope_kv_append(const __nv_bfloat16 *qkv, const unsigned long long *mask,
                                                        int mask_words, pia_slots_t sl,
                                                        int hq, int hkv, int hd, const __nv_bfloat16 *cos_t,
                                                        const __nv_bfloat16 *sin_t, int max_pos, __nv_bfloat16 *q_out,
                                                        __nv_bfloat16 *kc, __nv_bfloat16 *vc, int max_seq) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x;                       // activation row
  const int slot = i / sl.rows_per_slot, node = i % sl.rows_per_slot;
  const int n = sl.d_n[slot], P = sl.d_prefix_len[slot];
  const int pad_len = sl.d_pad_len ? sl.d_pad_len[slot] : 0;
  if (node >= n) return;
  kc += (long long)slot * sl.kv_slot_stride;
  vc += (long long)slot * sl.kv_slot_stride;
  int depth = -1;
  for (int w = 0; w < mask_words; ++w) depth += __popcll(mask[(long long)i * mask_words + w]);
  // rowsum(attention_mask) - 1 (modeling_llama.py:587): visible prefix keys [pad_len, P) + visible draft keys - 1
  int pos = (P > pad_len ? P - pad_len : 0) + depth;
  if (pos < 0) pos = 0;
  if (pos >= max_pos) pos = max_pos - 1;
  const int half = hd >> 1, cpr = hd >> 3;  // 16-byte chunks per head row
  const int row_elems = (hq + 2 * hkv) * hd;
  const __nv_bfloat16 *src = qkv + (long long)i * row_elems;
  const int cache_row = P + node;
  const int total = (hq + 2 * hkv) * cpr;
  for (int c = threadIdx.x; c < total; c += blockDim.x) {
    const int head = c / cpr, ch = c % cpr;
    const int d0 = ch * 8;
    Pack8 a; a.u = *reinterpret_cast<const uint4 *>(src + head * hd + d0);
    if (head < hq + hkv) {  // q or k: x*cos + rotate_half(x)*sin, every product/sum rounded to bf16 like eager torch
      const int dp = d0 < half ? d0 + half : d0 - half;
      Pack8 b; b.u = *reinterpret_cast<const uint4 *>(src + head * hd + dp);
      Pack8 cs, sn;
      const int f0 = d0 % half;
      cs.u = *reinterpret_cast<const uint4 *>(cos_t + (long long)pos * half + f0);
      sn.u = *reinterpret_cast<const uint4 *>(sin_t + (long long)pos * half + f0);
      Pack8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x = __bfloat162float(a.h[j]);
        const float r = d0 < half ? -__bfloat162float(b.h[j]) : __bfloat162float(b.h[j]);
        o.h[j] = __float2bfloat16_rn(bf(x * __bfloat162float(cs.h[j])) + bf(r * __bfloat162float(sn.h[j])));
      }
      if (head < hq) *reinterpret_cast<uint4 *>(q_out + ((long long)i * hq + head) * hd + d0) = o.u;
      else *reinterpret_cast<uint4 *>(kc + ((long long)(head - hq) * max_seq + cache_row) * hd + d0) = o.u;
    } else {
      *reinterpret_cast<uint4 *>(vc + ((long long)(head - hq - hkv) * max_seq + cache_row) * hd + d0) = a.u;
    }
  }
}

__global__ void __launch_bounds__(256) k_silu_mul(const __nv_bfloat16 *gu, int inter, __nv_bfloat16 *out) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.y;
  const int v = blockIdx.x * 256 + threadIdx.x;
  if (v * 8 >= inter) return;
  Pack8 g, u, o;
  g.u = *reinterpret_cast<const uint4 *>(gu + (long long)row * 2 * inter + v * 8);
  u.u = *reinterpret_cast<const uint4 *>(gu + (long long)row * 2 * inter + inter + v * 8);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float x = __bfloat162float(g.h[j]);
    const float s = bf(x / (1.f + expf(-x)));
    o.h[j] = __float2bfloat16_rn(s * __bfloat162float(u.h[j]));
  }
  *reinterpret_cast<uint4 *>(out + (long long)row * inter + v * 8) = o.u;
}

__global__ void __launch_bounds__(256) k_embed_gather(const __nv_bfloat16 *table, const int *ids, const int *d_n,
                                                      int hidden, __nv_bfloat16 *out) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  const int n = *d_n;
  const int nvec = hidden >> 3;
  uint4 *dst = reinterpret_cast<uint4 *>(out + (long long)row * hidden);
  if (row >= n) { for (int v = threadIdx.x; v < nvec; v += 256) dst[v] = make_uint4(0, 0, 0, 0); return; }
  const uint4 *src = reinterpret_cast<const uint4 *>(table + (long long)ids[row] * hidden);
  for (int v = threadIdx.x; v < nvec; v += 256) dst[v] = src[v];
}

}  // namespace fused
}  // namespace pia

using namespace pia;
using namespace pia::fused;

extern "C" int pia_rmsnorm(const void *d_x, const void *d_residual_in, const void *d_weight, float eps, int rows,
                           int hidden, void *d_residual_out, void *d_y, void *stream) {
  PIA_REQUIRE(d_x && d_weight && d_y && rows > 0 && hidden > 0 && hidden % 8 == 0, "bad rmsnorm arguments");
  PIA_CUDA_CHECK(launch_kernel(k_rmsnorm, dim3(rows), dim3(512), 0, (cudaStream_t)stream, (const __nv_bfloat16 *)d_x,
                              (const float *)nullptr, 0, 0ll, (const __nv_bfloat16 *)d_residual_in,
                              (const __nv_bfloat16 *)d_weight, eps, hidden, (__nv_bfloat16 *)d_residual_out,
                              (__nv_bfloat16 *)d_y));
  count_launch();
  return PIA_OK;
}

extern "C" int pia_rmsnorm_partials(const float *d_x_parts, int n_parts, int64_t part_stride, const void *d_residual_in,
                                    const void *d_weight, float eps, int rows, int hidden, void *d_residual_out,
                                    void *d_y, void *stream) {
  PIA_REQUIRE(d_x_parts && n_parts >= 1 && d_weight && d_y && rows > 0 && hidden > 0 && hidden % 8 == 0, "bad rmsnorm arguments");
  PIA_CUDA_CHECK(launch_kernel(k_rmsnorm, dim3(rows), dim3(512), 0, (cudaStream_t)stream,
                              (const __nv_bfloat16 *)nullptr, d_x_parts, n_parts, (long long)part_stride,
                              (const __nv_bfloat16 *)d_residual_in, (const __nv_bfloat16 *)d_weight, eps, hidden,
                              (__nv_bfloat16 *)d_residual_out, (__nv_bfloat16 *)d_y));
  count_launch();
  return PIA_OK;
}

extern "C" int pia_rope_kv_append(const void *d_qkv, const uint64_t *d_mask, int mask_words, const pia_slots_t *slots,
                                  int n_q_heads, int n_kv_heads, int head_dim, const void *d_cos, const void *d_sin,
                                  int max_pos, void *d_q_out, void *d_k_cache_layer, void *d_v_cache_layer, int max_seq,
                                  void *stream) {
  PIA_REQUIRE(d_qkv && d_mask && slots && slots->d_n && slots->d_prefix_len && d_cos && d_sin && d_q_out &&
                  d_k_cache_layer && d_v_cache_layer, "null argument");
  PIA_REQUIRE(slots->batch >= 1 && slots->rows_per_slot >= 1 && slots->kv_slot_stride >= 0, "bad slot table");
  PIA_REQUIRE(head_dim % 16 == 0 && mask_words >= 1 && mask_words <= 2, "bad rope arguments");
  PIA_CUDA_CHECK(launch_kernel(k_rope_kv_append, dim3(slots->batch * slots->rows_per_slot), dim3(256), 0,
                              (cudaStream_t)stream, (const __nv_bfloat16 *)d_qkv, (const unsigned long long *)d_mask,
                              mask_words, *slots, n_q_heads, n_kv_heads, head_dim, (const __nv_bfloat16 *)d_cos,
                              (const __nv_bfloat16 *)d_sin, max_pos, (__nv_bfloat16 *)d_q_out,
                              (__nv_bfloat16 *)d_k_cache_layer, (__nv_bfloat16 *)d_v_cache_layer, max_seq));
  count_launch();
  return PIA_OK;
}

extern "C" int pia_silu_mul(const void *d_gate_up, int rows, int inter, void *d_out, void *stream) {
  PIA_REQUIRE(d_gate_up && d_out && rows > 0 && inter > 0 && inter % 8 == 0, "bad silu_mul arguments");
  dim3 grid((inter / 8 + 255) / 256, rows);
  PIA_CUDA_CHECK(launch_kernel(k_silu_mul, grid, dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16 *)d_gate_up,
                              inter, (__nv_bfloat16 *)d_out));
  count_launch();
  return PIA_OK;
}

// MoE combine (mixtral/modeling_mixtral.py:734-759 restated densely): out[t] = sum over experts e, in expert-index
// order, of ye[e][t] * w[t][e]; the product and every partial sum are rounded to bf16 like the eager bf16 loop
// (`final_hidden_states.index_add_`), w = 0 for the experts a token did not select.  grid = rows, thread = 8 columns.
__global__ void __launch_bounds__(256) k_moe_combine(const __nv_bfloat16 *ye, const __nv_bfloat16 *w, int n_exp, int rows_cap,
                                                     int hidden, __nv_bfloat16 *out) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x;
  for (int v = threadIdx.x; v * 8 < hidden; v += blockDim.x) {
    Pack8 acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.h[j] = __float2bfloat16_rn(0.f);
    for (int e = 0; e < n_exp; ++e) {
      const float we = __bfloat162float(w[(long long)t * n_exp + e]);
      Pack8 y;
      y.u = *reinterpret_cast<const uint4 *>(ye + ((long long)e * rows_cap + t) * hidden + v * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        acc.h[j] = __float2bfloat16_rn(__bfloat162float(acc.h[j]) + bf(__bfloat162float(y.h[j]) * we));
    }
    *reinterpret_cast<uint4 *>(out + (long long)t * hidden + v * 8) = acc.u;
  }
}

extern "C" int pia_moe_combine(const void *d_expert_out, const void *d_weights, int n_experts, int rows, int rows_cap,
                               int hidden, void *d_out, void *stream) {
  PIA_REQUIRE(d_expert_out && d_weights && d_out && n_experts > 0 && rows > 0 && rows <= rows_cap && hidden % 8 == 0,
              "bad combine arguments");
  PIA_CUDA_CHECK(launch_kernel(k_moe_combine, dim3(rows), dim3(256), 0, (cudaStream_t)stream,
                              (const __nv_bfloat16 *)d_expert_out, (const __nv_bfloat16 *)d_weights, n_experts, rows_cap,
                              hidden, (__nv_bfloat16 *)d_out));
  count_launch();
  return PIA_OK;
}

// MoE router (mixtral/modeling_mixtral.py:721-727): router_logits = gate(hidden) (a bf16 Linear: fp32 accumulation,
// one bf16 rounding), softmax in fp32, top-k, renormalise, cast to bf16; written DENSE: w[t][e] = routing weight or 0 for
// the experts token t did not select (what k_moe_combine and the dense-over-experts verify path consume).
// grid = rows; one warp per expert dot product (round robin), thread 0 does the 8-way softmax / top-k.
__global__ void __launch_bounds__(256) k_moe_router(const __nv_bfloat16 *y, const __nv_bfloat16 *gate_w, int hidden,
                                                    int n_exp, int top_k, __nv_bfloat16 *dense) {
  __shared__ float s_logit[64];
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16 *yr = y + (long long)t * hidden;
  for (int e = warp; e < n_exp; e += 8) {
    const __nv_bfloat16 *wr = gate_w + (long long)e * hidden;
    float acc = 0.f;
    for (int v = lane; v * 8 < hidden; v += 32) {
      Pack8 a, w;
      a.u = *reinterpret_cast<const uint4 *>(yr + v * 8);
      w.u = *reinterpret_cast<const uint4 *>(wr + v * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += __bfloat162float(a.h[j]) * __bfloat162float(w.h[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if (lane == 0) s_logit[e] = bf(acc);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = -INFINITY;
    for (int e = 0; e < n_exp; ++e) mx = fmaxf(mx, s_logit[e]);
    float den = 0.f, p[64];
    for (int e = 0; e < n_exp; ++e) { p[e] = expf(s_logit[e] - mx); den += p[e]; }
    for (int e = 0; e < n_exp; ++e) p[e] /= den;
    unsigned long long chosen = 0ull;
    float sum = 0.f;
    for (int k = 0; k < top_k; ++k) {   // largest first, lowest index on ties
      int best = -1;
      for (int e = 0; e < n_exp; ++e)
        if (!((chosen >> e) & 1ull) && (best < 0 || p[e] > p[best])) best = e;
      chosen |= 1ull << best;
      sum += p[best];
    }
    for (int e = 0; e < n_exp; ++e)
      dense[(long long)t * n_exp + e] = __float2bfloat16_rn(((chosen >> e) & 1ull) ? p[e] / sum : 0.f);
  }
}

extern "C" int pia_moe_router(const void *d_y, const void *d_gate_weight, int rows, int hidden, int n_experts, int top_k,
                              void *d_dense_out, void *stream) {
  PIA_REQUIRE(d_y && d_gate_weight && d_dense_out && rows > 0 && hidden % 8 == 0 && n_experts >= 1 && n_experts <= 64 &&
                  top_k >= 1 && top_k <= n_experts, "bad router arguments");
  PIA_CUDA_CHECK(launch_kernel(k_moe_router, dim3(rows), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16 *)d_y,
                              (const __nv_bfloat16 *)d_gate_weight, hidden, n_experts, top_k, (__nv_bfloat16 *)d_dense_out));
  count_launch();
  return PIA_OK;
}

// One warp per SM walks the ranges chunk by chunk (the bulk-prefetch issue rate of a single SM's TMA unit is only a
// few hundred GB/s, so the chunks are dealt round-robin over the whole grid); bytes_per_ns paces the grid against
// %globaltimer so that the demand loads of the kernels running beside it (attention's KV tiles) are not queued behind
// tens of MB of prefetch.  No shared memory, 32 threads: fits next to any resident CTA.
__global__ void __launch_bounds__(32) k_l2_prefetch(const char *base, long long n_ranges, long long stride,
                                                     long long range_bytes, int chunk, float bytes_per_ns) {
  const long long cpr = (range_bytes + chunk - 1) / chunk;
  const long long total = cpr * n_ranges;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (long long c = (long long)threadIdx.x * gridDim.x + blockIdx.x; c < total; c += (long long)blockDim.x * gridDim.x) {
    const long long k = c / n_ranges, r = c % n_ranges;
    const long long off = k * chunk;
    const long long left = range_bytes - off;
    const unsigned sz = (unsigned)(left < chunk ? left : chunk);
    if (bytes_per_ns > 0.f) {
      const unsigned long long due = t0 + (unsigned long long)((float)(c * chunk) / bytes_per_ns);
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      while (now < due) {
        __nanosleep(64);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      }
    }
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + r * stride + off), "r"(sz) : "memory");
  }
}

extern "C" int pia_l2_prefetch(const void *d_base, int64_t n_ranges, int64_t stride_bytes, int64_t range_bytes,
                               float gbytes_per_s, void *stream) {
  PIA_REQUIRE(d_base && n_ranges > 0 && range_bytes > 0 && range_bytes % 16 == 0 && ((uintptr_t)d_base & 15) == 0 &&
                  (n_ranges == 1 || (stride_bytes >= range_bytes && stride_bytes % 16 == 0)) && gbytes_per_s >= 0.f,
              "bad prefetch arguments");
  const int chunk = range_bytes < 16384 ? (int)range_bytes : 16384;
  // plain launch (no programmatic dependency): the kernel reads nothing its predecessors write
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    PIA_CUDA_CHECK(cudaGetDevice(&dev));
    PIA_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  k_l2_prefetch<<<n_sm, 32, 0, (cudaStream_t)stream>>>((const char *)d_base, (long long)n_ranges, (long long)stride_bytes,
                                                     (long long)range_bytes, chunk, gbytes_per_s);
  PIA_LAUNCH_CHECK();
  return PIA_OK;
}

extern "C" int pia_embed_gather(const void *d_table, const int32_t *d_ids, const int32_t *d_n, int rows, int hidden,
                                void *d_out, void *stream) {
  PIA_REQUIRE(d_table && d_ids && d_n && d_out && rows > 0 && hidden % 8 == 0, "bad embed arguments");
  PIA_CUDA_CHECK(launch_kernel(k_embed_gather, dim3(rows), dim3(256), 0, (cudaStream_t)stream,
                              (const __nv_bfloat16 *)d_table, (const int *)d_ids, (const int *)d_n, hidden,
                              (__nv_bfloat16 *)d_out));
  count_launch();
  return PIA_OK;
}
