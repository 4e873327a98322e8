// Row-level helpers shared by the backward kernels: table lookup by global key, AdamW row arithmetic.
#pragma once
#include "tgr_common.cuh"

namespace tgr {

struct RowParams {
  float* w[TGR_MAX_TABLES];
  float* m[TGR_MAX_TABLES];
  float* v[TGR_MAX_TABLES];
  float* grad[TGR_MAX_TABLES];
  uint32_t key_base[TGR_MAX_TABLES + 1];
  int32_t n_tables;
  int32_t H4;
  tgr_adam_t adam;
  const tgr_adam_t* adam_dev;   // != NULL: the hyper-parameters are read from device memory when the kernel runs (graph replay)
};

__device__ __forceinline__ float adam_elem(float& w, float& m, float& v, float g, const tgr_adam_t& a) {
  // torch/optim/adam.py _single_tensor_adam with decoupled weight decay, rounding for rounding as the CPU
  // kernels evaluate it (probed against torch 2.11 CPU: lerp_ and addcmul_ fuse their last multiply-add):
  //   param.mul_(1 - lr*wd)                                   w = w * decay
  //   exp_avg.lerp_(grad, 1-beta1)                            m = fma(1-b1, g - m, m)
  //   exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)    v = fma((1-b2)*g, g, v*b2)
  //   denom = exp_avg_sq.sqrt() / bc2_sqrt + eps
  //   param.addcdiv_(exp_avg, denom, value=-step_size)        w = w + (-step_size*m)/denom
  w = __fmul_rn(w, a.decay);
  m = __fmaf_rn(a.one_minus_beta1, __fsub_rn(g, m), m);
  v = __fmaf_rn(__fmul_rn(a.one_minus_beta2, g), g, __fmul_rn(v, a.beta2));
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), a.bc2_sqrt), a.eps);
  w = __fadd_rn(w, __fdiv_rn(__fmul_rn(-a.step_size, m), denom));
  return w;
}

__device__ __forceinline__ void adam_row4(float4* wp, float4* mp, float4* vp, float4 g, const tgr_adam_t& a) {
  float4 w = *wp, m = *mp, v = *vp;
  g.x *= a.grad_scale; g.y *= a.grad_scale; g.z *= a.grad_scale; g.w *= a.grad_scale;
  adam_elem(w.x, m.x, v.x, g.x, a);
  adam_elem(w.y, m.y, v.y, g.y, a);
  adam_elem(w.z, m.z, v.z, g.z, a);
  adam_elem(w.w, m.w, v.w, g.w, a);
  *wp = w; *mp = m; *vp = v;
}

__device__ __forceinline__ int find_table(const uint32_t* key_base, int n_tables, uint32_t key) {
  int lo = 0, hi = n_tables;  // key_base[lo] <= key < key_base[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (key >= key_base[mid]) lo = mid; else hi = mid;
  }
  return lo;
}

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

inline int fill_row_params(RowParams& rp, const tgr_table_t* tables, int n_tables, int H) {
  TGR_REQUIRE(tables && n_tables > 0 && n_tables <= TGR_MAX_TABLES, "bad table array");
  TGR_REQUIRE(H > 0 && H % 4 == 0, "bad H=%d", H);
  rp.n_tables = n_tables;
  rp.H4 = H / 4;
  for (int t = 0; t < n_tables; ++t) {
    rp.w[t] = tables[t].weight;
    rp.m[t] = tables[t].exp_avg;
    rp.v[t] = tables[t].exp_avg_sq;
    rp.grad[t] = tables[t].grad;
    rp.key_base[t] = (uint32_t)tables[t].key_base;
    if (t) TGR_REQUIRE(tables[t].key_base == tables[t - 1].key_base + tables[t - 1].rows, "key bases must be cumulative");
  }
  rp.key_base[n_tables] = (uint32_t)(tables[n_tables - 1].key_base + tables[n_tables - 1].rows);
  return 0;
}


}  // namespace tgr
