// Small vector-type helpers shared by the SpMM kernels (128-bit accesses of right-hand-side rows).
#pragma once

#include "common.cuh"

namespace mgp {

template <typename T, int VEC>
struct Vec;
template <typename T>
struct Vec<T, 1> {
  T v[1];
};
template <>
struct alignas(16) Vec<float, 4> {
  float v[4];
};
template <>
struct alignas(16) Vec<double, 2> {
  double v[2];
};

template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> ldg_vec(const T* p) {
  Vec<T, VEC> r;
  if constexpr (VEC == 1) {
    r.v[0] = __ldg(p);
  } else if constexpr (sizeof(T) == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    const double2 t = __ldg(reinterpret_cast<const double2*>(p));
    r.v[0] = t.x; r.v[1] = t.y;
  }
  return r;
}

template <typename T, int VEC>
__device__ __forceinline__ void st_vec(T* p, const Vec<T, VEC>& r) {
  if constexpr (VEC == 1) {
    *p = r.v[0];
  } else if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else {
    *reinterpret_cast<double2*>(p) = make_double2(r.v[0], r.v[1]);
  }
}


// block-level column reduction + deterministic last-block reduction used by the dot-product epilogues
template <typename T, int VEC, int LPN, int CWMAX, int BLOCK>
__device__ __forceinline__ bool spmm_dot_epilogue(T (&dsum)[VEC], int cw, int c0, T* partials, unsigned int* counter,
                                                  T* dot_out) {
  __shared__ T sm_dot[BLOCK / 32][CWMAX];
  __shared__ T red_dot[BLOCK];
  const int tid = threadIdx.x, lane = tid & 31;
#pragma unroll
  for (int v = 0; v < VEC; ++v) dsum[v] = subwarp_sum(dsum[v], LPN, 32);
  if (lane < LPN) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) sm_dot[tid >> 5][lane * VEC + v] = dsum[v];
  }
  __syncthreads();
  if (tid < cw) {
    T s = T(0);
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) s += sm_dot[w][tid];
    partials[(int64_t)blockIdx.x * cw + tid] = s;
  }
  const bool is_last_block = last_block_ticket_writers(counter, tid < cw);
  if (is_last_block) {
    const int c = tid % cw;
    const int lanes_per_col = BLOCK / cw;  // cw <= 32
    const int r = tid / cw;
    T s = T(0);
    if (r < lanes_per_col)
      for (int b = r; b < (int)gridDim.x; b += lanes_per_col) s += __ldcg(partials + (int64_t)b * cw + c);
    red_dot[tid] = (r < lanes_per_col) ? s : T(0);
    __syncthreads();
    if (tid < cw) {
      T t = T(0);
      for (int rr = 0; rr < lanes_per_col; ++rr) t += red_dot[rr * cw + tid];
      dot_out[c0 + tid] = t;
    }
  }
  return is_last_block;     // block-uniform: true in the block that wrote dot_out
}

}  // namespace mgp
