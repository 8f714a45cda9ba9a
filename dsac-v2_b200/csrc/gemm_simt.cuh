// fp32 grouped GEMM for the dense layers of the DSAC-T update (DSACT_GEMM_FP32).
//
// One launch processes up to MAXG independent problems (the twin critics, their
// targets, the actor pass ...), each
//     C[M,N] (op)= sum over <=2 K-segments of  A_s[M,K_s] * B_s[K_s,N]
// with the operand orientation as a template parameter so that the three shapes
// of a linear layer map onto one kernel without materialising a transpose:
//     forward  y = x W^T      : A k-contiguous, B k-contiguous   (AK=1, BK=1)
//     dgrad    dx = dy W      : A k-contiguous, B n-contiguous   (AK=1, BK=0)
//     wgrad    dW = dy^T x    : A m-contiguous, B n-contiguous   (AK=0, BK=0), split-K + atomics
// The second K segment is how cat(obs, act) (reference networks/mlp.py:123) is
// consumed without ever building the concatenation.
//
// Epilogues fuse bias + activation (+ pre-activation store), activation
// derivative (+ bias-gradient column sums), or the split-K atomic accumulate.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dsact {

enum { ACT_LINEAR = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_TANH = 3, ACT_SIGMOID = 4, ACT_ELU = 5, ACT_SELU = 6 };

__device__ __forceinline__ float act_fwd(float z, int a) {
  switch (a) {
    case ACT_GELU: return 0.5f * z * (1.0f + erff(z * 0.70710678118654752f));  // nn.GELU() exact
    case ACT_RELU: return fmaxf(z, 0.0f);
    case ACT_TANH: return tanhf(z);
    case ACT_SIGMOID: return 1.0f / (1.0f + expf(-z));
    case ACT_ELU: return z > 0.0f ? z : expm1f(z);
    case ACT_SELU: return 1.0507009873554805f * (z > 0.0f ? z : 1.6732632423543772f * expm1f(z));
    default: return z;
  }
}

__device__ __forceinline__ float act_bwd(float z, int a) {
  switch (a) {
    case ACT_GELU:
      return 0.5f * (1.0f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * expf(-0.5f * z * z);
    case ACT_RELU: return z > 0.0f ? 1.0f : 0.0f;
    case ACT_TANH: { float t = tanhf(z); return 1.0f - t * t; }
    case ACT_SIGMOID: { float s = 1.0f / (1.0f + expf(-z)); return s * (1.0f - s); }
    case ACT_ELU: return z > 0.0f ? 1.0f : expf(z);
    case ACT_SELU: return 1.0507009873554805f * (z > 0.0f ? 1.0f : 1.6732632423543772f * expf(z));
    default: return 1.0f;
  }
}

enum { EPI_STORE = 0, EPI_BIAS_ACT = 1, EPI_DACT = 2, EPI_ATOMIC = 3 };

constexpr int MAXG = 8;
constexpr int KT = 16;  // k-tile depth

struct GemmProb {
  const float* A[2];
  const float* B[2];
  float* C;
  const float* bias;   // [N] or null           (EPI_STORE / EPI_BIAS_ACT)
  float* Zout;         // pre-activation or null (EPI_BIAS_ACT), leading dim ldc
  const float* Zin;    // pre-activation         (EPI_DACT), leading dim ldz
  float* colsum;       // [N] += column sums of the result or null (EPI_DACT)
  int lda[2], ldb[2], K[2];
  int M, N, ldc, ldz;
  int epi, act;
  int tiles_m, tiles_n, ksplit, tile_start;
};

struct GemmGroup {
  int n;
  GemmProb p[MAXG];
};

__device__ __forceinline__ float4 load4(const float* __restrict__ p, int valid, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid >= 4 && vec) {
    v = __ldg(reinterpret_cast<const float4*>(p));
  } else {
    if (valid > 0) v.x = __ldg(p);
    if (valid > 1) v.y = __ldg(p + 1);
    if (valid > 2) v.z = __ldg(p + 2);
    if (valid > 3) v.w = __ldg(p + 3);
  }
  return v;
}

__device__ __forceinline__ bool vec_ok(const float* p, int ld) {
  return ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ((ld & 3) == 0);
}

// 256 threads as 16x16; each owns a (BM/16)x(BN/16) micro-tile split in 4-wide groups so that
// the shared-memory reads of a quarter-warp are 128 contiguous bytes.
template <int BM, int BN, bool AK, bool BK>
__global__ void __launch_bounds__(256) gemm_kernel(const __grid_constant__ GemmGroup g) {
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int GM = TM / 4, GN = TN / 4;
  constexpr int LA = (BM * KT / 4) / 256;  // float4 loads per thread for the A tile
  constexpr int LB = (BN * KT / 4) / 256;
  static_assert(LA >= 1 && LB >= 1, "tile too small");

  __shared__ __align__(16) float As[2][KT][BM + 4];
  __shared__ __align__(16) float Bs[2][KT][BN + 4];

  asm volatile("griddepcontrol.wait;" ::: "memory");           // programmatic dependent launch (see kernels.cuh)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int t = threadIdx.x;
  int pi = 0;
#pragma unroll
  for (int i = 1; i < MAXG; ++i)
    if (i < g.n && (int)blockIdx.x >= g.p[i].tile_start) pi = i;
  const GemmProb& P = g.p[pi];
  int local = blockIdx.x - P.tile_start;
  const int tiles_mn = P.tiles_m * P.tiles_n;
  const int ks = local / tiles_mn;
  local -= ks * tiles_mn;
  const int m0 = (local / P.tiles_n) * BM, n0 = (local % P.tiles_n) * BN;
  const int M = P.M, N = P.N;

  // k-tile schedule: segment 0 (optionally split across CTAs), then segment 1
  const int nt0 = (P.K[0] + KT - 1) / KT, nt1 = (P.K[1] + KT - 1) / KT;
  int t_begin = 0, t_end = nt0 + nt1;
  if (P.ksplit > 1) {
    const int per = (nt0 + P.ksplit - 1) / P.ksplit;
    t_begin = ks * per;
    t_end = min(nt0, t_begin + per);
  }

  const bool va0 = vec_ok(P.A[0], P.lda[0]), va1 = P.K[1] ? vec_ok(P.A[1], P.lda[1]) : false;
  const bool vb0 = vec_ok(P.B[0], P.ldb[0]), vb1 = P.K[1] ? vec_ok(P.B[1], P.ldb[1]) : false;

  float4 ra[LA], rb[LB];

  auto fetch = [&](int kt) {
    const int s = kt >= nt0 ? 1 : 0;
    const int k0 = (s ? kt - nt0 : kt) * KT;
    const float* __restrict__ A = P.A[s];
    const float* __restrict__ B = P.B[s];
    const int lda = P.lda[s], ldb = P.ldb[s], K = P.K[s];
    const bool va = s ? va1 : va0, vb = s ? vb1 : vb0;
#pragma unroll
    for (int r = 0; r < LA; ++r) {
      const int id = t + r * 256;
      if (AK) {
        const int row = id % BM, kq = id / BM;
        const int m = m0 + row, k = k0 + kq * 4;
        const int valid = m < M ? min(4, K - k) : 0;
        ra[r] = load4(A + (size_t)m * lda + k, valid, va);
      } else {
        const int kk = id / (BM / 4), mq = id % (BM / 4);
        const int k = k0 + kk, m = m0 + mq * 4;
        const int valid = k < K ? min(4, M - m) : 0;
        ra[r] = load4(A + (size_t)k * lda + m, valid, va);
      }
    }
#pragma unroll
    for (int r = 0; r < LB; ++r) {
      const int id = t + r * 256;
      if (BK) {
        const int row = id % BN, kq = id / BN;
        const int n = n0 + row, k = k0 + kq * 4;
        const int valid = n < N ? min(4, K - k) : 0;
        rb[r] = load4(B + (size_t)n * ldb + k, valid, vb);
      } else {
        const int kk = id / (BN / 4), nq = id % (BN / 4);
        const int k = k0 + kk, n = n0 + nq * 4;
        const int valid = k < K ? min(4, N - n) : 0;
        rb[r] = load4(B + (size_t)k * ldb + n, valid, vb);
      }
    }
  };

  auto stash = [&](int buf) {
#pragma unroll
    for (int r = 0; r < LA; ++r) {
      const int id = t + r * 256;
      if (AK) {
        const int row = id % BM, kq = id / BM;
        As[buf][kq * 4 + 0][row] = ra[r].x;
        As[buf][kq * 4 + 1][row] = ra[r].y;
        As[buf][kq * 4 + 2][row] = ra[r].z;
        As[buf][kq * 4 + 3][row] = ra[r].w;
      } else {
        const int kk = id / (BM / 4), mq = id % (BM / 4);
        *reinterpret_cast<float4*>(&As[buf][kk][mq * 4]) = ra[r];
      }
    }
#pragma unroll
    for (int r = 0; r < LB; ++r) {
      const int id = t + r * 256;
      if (BK) {
        const int row = id % BN, kq = id / BN;
        Bs[buf][kq * 4 + 0][row] = rb[r].x;
        Bs[buf][kq * 4 + 1][row] = rb[r].y;
        Bs[buf][kq * 4 + 2][row] = rb[r].z;
        Bs[buf][kq * 4 + 3][row] = rb[r].w;
      } else {
        const int kk = id / (BN / 4), nq = id % (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][kk][nq * 4]) = rb[r];
      }
    }
  };

  const int tx = t % 16, ty = t / 16;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  if (t_begin < t_end) {
    fetch(t_begin);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int kt = t_begin; kt < t_end; ++kt) {
      if (kt + 1 < t_end) fetch(kt + 1);
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        float a[TM], b[TN];
#pragma unroll
        for (int gi = 0; gi < GM; ++gi) {
          const float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][gi * (BM / GM) + ty * 4]);
          a[gi * 4 + 0] = v.x; a[gi * 4 + 1] = v.y; a[gi * 4 + 2] = v.z; a[gi * 4 + 3] = v.w;
        }
#pragma unroll
        for (int gj = 0; gj < GN; ++gj) {
          const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][gj * (BN / GN) + tx * 4]);
          b[gj * 4 + 0] = v.x; b[gj * 4 + 1] = v.y; b[gj * 4 + 2] = v.z; b[gj * 4 + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (kt + 1 < t_end) {
        stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
      }
    }
  }

  // ---- epilogue
  const int epi = P.epi, act = P.act;
  float* __restrict__ C = P.C;
  const int ldc = P.ldc;
  const bool vc = vec_ok(C, ldc) && epi != EPI_ATOMIC;
  float csum[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) csum[j] = 0.f;

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i / 4) * (BM / GM) + ty * 4 + (i % 4);
    if (m >= M) continue;
#pragma unroll
    for (int gj = 0; gj < GN; ++gj) {
      const int n = n0 + gj * (BN / GN) + tx * 4;
      if (n >= N) continue;
      const int valid = min(4, N - n);
      float v[4], z[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[i][gj * 4 + j];
      if (epi == EPI_ATOMIC) {
        for (int j = 0; j < valid; ++j) atomicAdd(C + (size_t)m * ldc + n + j, v[j]);
        continue;
      }
      if (epi == EPI_STORE || epi == EPI_BIAS_ACT) {
        if (P.bias)
          for (int j = 0; j < valid; ++j) v[j] += __ldg(P.bias + n + j);
        if (epi == EPI_BIAS_ACT) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { z[j] = v[j]; v[j] = act_fwd(v[j], act); }
          if (P.Zout) {
            float* zp = P.Zout + (size_t)m * ldc + n;
            if (valid == 4 && vc && vec_ok(P.Zout, ldc)) *reinterpret_cast<float4*>(zp) = make_float4(z[0], z[1], z[2], z[3]);
            else for (int j = 0; j < valid; ++j) zp[j] = z[j];
          }
        }
      } else {  // EPI_DACT
        const float* zp = P.Zin + (size_t)m * P.ldz + n;
        for (int j = 0; j < valid; ++j) {
          v[j] *= act_bwd(__ldg(zp + j), act);
          csum[gj * 4 + j] += v[j];
        }
      }
      float* cp = C + (size_t)m * ldc + n;
      if (valid == 4 && vc) *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
      else for (int j = 0; j < valid; ++j) cp[j] = v[j];
    }
  }

  if (epi == EPI_DACT && P.colsum) {  // bias gradient: column sums of this tile
    __syncthreads();
    float* red = &As[0][0][0];  // >= 16*BN floats
#pragma unroll
    for (int j = 0; j < TN; ++j) red[ty * BN + (j / 4) * (BN / GN) + tx * 4 + (j % 4)] = csum[j];
    __syncthreads();
    if (t < BN && n0 + t < N) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) s += red[r * BN + t];
      atomicAdd(P.colsum + n0 + t, s);
    }
  }
}

}  // namespace dsact
