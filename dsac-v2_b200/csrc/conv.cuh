// Convolution kernels of the CNN approximators (BASELINE config 5; reference networks/cnn.py:30-53: Conv2d + ReLU per
// layer, no padding, square kernels, NCHW fp32).  Direct convolutions in fp32 (the reference's arithmetic up to summation
// order); the dense heads behind the encoder run through the grouped GEMM kernels of gemm_simt.cuh, and so does the weight
// gradient of a layer whose window covers its whole input (the last layer of the type_2 encoder: a plain linear layer).
//
// Every kernel indexes ROWS = (sample, position) pairs flattened over the batch, so that the deep layers (5x5, 3x3, 1x1
// maps) fill their blocks as well as the first ones do, and holds eight channels per thread: one loaded value feeds eight
// FMAs against weights staged in shared memory with the eight channels of a tap adjacent (two 16-byte broadcasts).
//
//   conv_fwd8_kernel    y = relu(conv(x, w) + b)            thread = R (1, 2, 4) output positions x 8 output channels
//   conv_dgrad8_kernel  dx = convT(dy, w) (.) [x > 0]       thread = 1 input position x 8 input channels
//   conv_wgrad_kernel   dw += corr(x, dy), db += sum(dy)    block = (ci, 4 or 8 output channels, slab of rows); per-thread
//                                                           K*K x 4 (8) accumulators, block reduction, one atomicAdd per weight
//   conv_fwd_kernel / conv_dgrad_kernel                     one channel per thread (channel counts not divisible by 8)
// dy is the gradient w.r.t. a layer's pre-activation (the caller masked the top one; conv_dgrad masks what it hands down
// by the ReLU of the layer below: x is that layer's output, x > 0 <=> the unit was active).
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace dsact {

struct ConvShape {
  int B, Cin, Hin, Win, Cout, K, S, Hout, Wout;
};

// grid: (ceil(B*Hout*Wout / 128), Cout); block 128; dynamic smem Cin*K*K floats
__global__ void conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                float* __restrict__ y, const ConvShape s) {
  pdl_sync();
  extern __shared__ float wsm[];   // [Cin][K][K] of this block's output channel
  const int co = blockIdx.y;
  const int nw = s.Cin * s.K * s.K;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) wsm[i] = w[(size_t)co * nw + i];
  __syncthreads();
  const int npos = s.Hout * s.Wout;
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= (long long)s.B * npos) return;
  const int b = (int)(r / npos), p = (int)(r - (long long)b * npos);
  const int oy = p / s.Wout, ox = p - oy * s.Wout;
  const float* xb = x + (size_t)b * s.Cin * s.Hin * s.Win;
  float acc = bias[co];
  for (int ci = 0; ci < s.Cin; ++ci) {
    const float* xc = xb + (size_t)ci * s.Hin * s.Win + (size_t)(oy * s.S) * s.Win + ox * s.S;
    const float* wc = wsm + ci * s.K * s.K;
    for (int ky = 0; ky < s.K; ++ky)
      for (int kx = 0; kx < s.K; ++kx) acc = fmaf(xc[ky * s.Win + kx], wc[ky * s.K + kx], acc);
  }
  y[((size_t)b * s.Cout + co) * npos + p] = fmaxf(acc, 0.f);
}

// R output positions x 8 output channels per thread; K compile-time so that the tap loops unroll and the R*K*K loads of
// one input channel are issued together.  grid: (ceil(B*Hout*Wout / (128*R)), Cout / 8); block 128; dynamic smem
// 8 * Cin * K * K floats.
template <int K, int R>
__global__ void __launch_bounds__(128) conv_fwd8_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y, const ConvShape s) {
  pdl_sync();
  extern __shared__ float wsm[];   // [Cin*K*K][8]
  constexpr int KK = K * K;
  const int co0 = blockIdx.y * 8;
  const int nw = s.Cin * KK;
  for (int i = threadIdx.x; i < 8 * nw; i += blockDim.x) {
    const int c = i / nw, tap = i - c * nw;
    wsm[tap * 8 + c] = w[(size_t)(co0 + c) * nw + tap];
  }
  __syncthreads();
  const int npos = s.Hout * s.Wout;
  const long long rows = (long long)s.B * npos;
  const float* xr[R];
  long long yoff[R];
  bool live[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    long long r = ((long long)blockIdx.x * R + i) * blockDim.x + threadIdx.x;
    live[i] = r < rows;
    if (!live[i]) r = 0;
    const int b = (int)(r / npos), p = (int)(r - (long long)b * npos);
    const int oy = p / s.Wout, ox = p - oy * s.Wout;
    xr[i] = x + (size_t)b * s.Cin * s.Hin * s.Win + (size_t)(oy * s.S) * s.Win + ox * s.S;
    yoff[i] = (long long)b * s.Cout * npos + p;
  }
  float acc[R][8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float bv = bias[co0 + c];
#pragma unroll
    for (int i = 0; i < R; ++i) acc[i][c] = bv;
  }
  const size_t plane = (size_t)s.Hin * s.Win;
  for (int ci = 0; ci < s.Cin; ++ci) {
    const float* wc = wsm + (size_t)ci * KK * 8;
    float v[R][KK];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int ky = 0; ky < K; ++ky)
#pragma unroll
        for (int kx = 0; kx < K; ++kx) v[i][ky * K + kx] = xr[i][ci * plane + ky * s.Win + kx];
#pragma unroll
    for (int t = 0; t < KK; ++t) {
      const float4 w0 = *reinterpret_cast<const float4*>(wc + t * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(wc + t * 8 + 4);
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const float u = v[i][t];
        acc[i][0] = fmaf(u, w0.x, acc[i][0]); acc[i][1] = fmaf(u, w0.y, acc[i][1]); acc[i][2] = fmaf(u, w0.z, acc[i][2]); acc[i][3] = fmaf(u, w0.w, acc[i][3]);
        acc[i][4] = fmaf(u, w1.x, acc[i][4]); acc[i][5] = fmaf(u, w1.y, acc[i][5]); acc[i][6] = fmaf(u, w1.z, acc[i][6]); acc[i][7] = fmaf(u, w1.w, acc[i][7]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < R; ++i)
    if (live[i]) {
#pragma unroll
      for (int c = 0; c < 8; ++c) y[yoff[i] + (size_t)(co0 + c) * npos] = fmaxf(acc[i][c], 0.f);
    }
}

// dL/dx masked by x > 0 when `mask_by_x`.  grid: (ceil(B*Hin*Win / 128), Cin); block 128
__global__ void conv_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, const float* __restrict__ x,
                                  float* __restrict__ dx, const ConvShape s, int mask_by_x) {
  pdl_sync();
  const int ci = blockIdx.y;
  const int npin = s.Hin * s.Win, npos = s.Hout * s.Wout;
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= (long long)s.B * npin) return;
  const int b = (int)(r / npin), p = (int)(r - (long long)b * npin);
  const int iy = p / s.Win, ix = p - iy * s.Win;
  const size_t xi = ((size_t)b * s.Cin + ci) * npin + p;
  if (mask_by_x && !(x[xi] > 0.f)) { dx[xi] = 0.f; return; }
  float acc = 0.f;
  for (int ky = 0; ky < s.K; ++ky) {
    const int ty = iy - ky;
    if (ty < 0 || ty % s.S) continue;
    const int oy = ty / s.S;
    if (oy >= s.Hout) continue;
    for (int kx = 0; kx < s.K; ++kx) {
      const int tx = ix - kx;
      if (tx < 0 || tx % s.S) continue;
      const int ox = tx / s.S;
      if (ox >= s.Wout) continue;
      const float* dyp = dy + (size_t)b * s.Cout * npos + (size_t)oy * s.Wout + ox;
      const float* wp = w + ((size_t)ci * s.K + ky) * s.K + kx;
      for (int co = 0; co < s.Cout; ++co) acc = fmaf(dyp[(size_t)co * npos], wp[(size_t)co * s.Cin * s.K * s.K], acc);
    }
  }
  dx[xi] = acc;
}

// Eight input channels per thread.  grid: (ceil(B*Hin*Win / 128), Cin / 8); block 128; dynamic smem Cout*K*K*8 floats
// (opt-in above 48 KB: the caller sets the attribute).
__global__ void __launch_bounds__(128) conv_dgrad8_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                          const float* __restrict__ x, float* __restrict__ dx, const ConvShape s,
                                                          int mask_by_x) {
  pdl_sync();
  extern __shared__ float wsm[];   // [Cout][K*K][8 input channels]
  const int ci0 = blockIdx.y * 8;
  const int KK = s.K * s.K;
  const int nw = s.Cout * KK;
  for (int i = threadIdx.x; i < 8 * nw; i += blockDim.x) {
    const int c = i / nw, rest = i - c * nw;          // rest = co*KK + tap
    const int co = rest / KK, tap = rest - co * KK;
    wsm[rest * 8 + c] = w[((size_t)co * s.Cin + ci0 + c) * KK + tap];
  }
  __syncthreads();
  const int npin = s.Hin * s.Win, npos = s.Hout * s.Wout;
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= (long long)s.B * npin) return;
  const int b = (int)(r / npin), p = (int)(r - (long long)b * npin);
  const int iy = p / s.Win, ix = p - iy * s.Win;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  const float* dyb = dy + (size_t)b * s.Cout * npos;
  // taps with (iy - ky) % S == 0: ky = iy % S, iy % S + S, ...  (every lane runs the same loop; only the tap differs)
  for (int ky = iy % s.S; ky < s.K && ky <= iy; ky += s.S) {
    const int oy = (iy - ky) / s.S;
    if (oy >= s.Hout) continue;
    for (int kx = ix % s.S; kx < s.K && kx <= ix; kx += s.S) {
      const int ox = (ix - kx) / s.S;
      if (ox >= s.Wout) continue;
      const float* dyp = dyb + (size_t)oy * s.Wout + ox;
      const float* wt = wsm + (size_t)(ky * s.K + kx) * 8;
#pragma unroll 8
      for (int co = 0; co < s.Cout; ++co) {
        const float g = dyp[(size_t)co * npos];
        const float4 w0 = *reinterpret_cast<const float4*>(wt + (size_t)co * KK * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wt + (size_t)co * KK * 8 + 4);
        acc[0] = fmaf(g, w0.x, acc[0]); acc[1] = fmaf(g, w0.y, acc[1]); acc[2] = fmaf(g, w0.z, acc[2]); acc[3] = fmaf(g, w0.w, acc[3]);
        acc[4] = fmaf(g, w1.x, acc[4]); acc[5] = fmaf(g, w1.y, acc[5]); acc[6] = fmaf(g, w1.z, acc[6]); acc[7] = fmaf(g, w1.w, acc[7]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const size_t xi = ((size_t)b * s.Cin + ci0 + c) * npin + p;
    dx[xi] = (mask_by_x && !(x[xi] > 0.f)) ? 0.f : acc[c];
  }
}

// dw[co][ci][ky][kx] += sum_{b,oy,ox} dy[b][co][oy][ox] * x[b][ci][oy*S+ky][ox*S+kx];  db[co] += sum dy (ci == 0 blocks).
// grid: (Cin, Cout / COB, slabs); block 256.  The caller cleared dw / db (begin_step_kernel); slabs split the rows.
template <int K, int COB>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                         float* __restrict__ dw, float* __restrict__ db, const ConvShape s) {
  pdl_sync();
  constexpr int KK = K * K, NA = KK * COB;
  const int ci = blockIdx.x, co0 = blockIdx.y * COB;
  float acc[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) acc[i] = 0.f;
  float bsum[COB];
#pragma unroll
  for (int c = 0; c < COB; ++c) bsum[c] = 0.f;
  const int npos = s.Hout * s.Wout;
  const long long total = (long long)s.B * npos;
  const long long per = (total + gridDim.z - 1) / gridDim.z;
  const long long r0 = per * blockIdx.z, r1 = r0 + per < total ? r0 + per : total;
  for (long long t = r0 + threadIdx.x; t < r1; t += blockDim.x) {
    const int b = (int)(t / npos), p = (int)(t - (long long)b * npos);
    const int oy = p / s.Wout, ox = p - oy * s.Wout;
    float g[COB];
#pragma unroll
    for (int c = 0; c < COB; ++c) { g[c] = dy[((size_t)b * s.Cout + co0 + c) * npos + p]; bsum[c] += g[c]; }
    const float* xp = x + ((size_t)b * s.Cin + ci) * s.Hin * s.Win + (size_t)(oy * s.S) * s.Win + ox * s.S;
#pragma unroll
    for (int ky = 0; ky < K; ++ky)
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const float v = xp[ky * s.Win + kx];
#pragma unroll
        for (int c = 0; c < COB; ++c) acc[(ky * K + kx) * COB + c] = fmaf(g[c], v, acc[(ky * K + kx) * COB + c]);
      }
  }
  __shared__ float red[(NA + COB) * 8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) red[i * 8 + wid] = v;
  }
#pragma unroll
  for (int c = 0; c < COB; ++c) {
    const float v = warp_sum(bsum[c]);
    if (lane == 0) red[(NA + c) * 8 + wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < NA + COB) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[threadIdx.x * 8 + k];
    if (threadIdx.x < NA) {
      const int tap = threadIdx.x / COB, c = threadIdx.x - tap * COB;
      atomicAdd(&dw[((size_t)(co0 + c) * s.Cin + ci) * KK + tap], v);
    } else if (ci == 0) {
      atomicAdd(&db[co0 + threadIdx.x - NA], v);
    }
  }
}

// out[n] += sum_m a[m][n]   (bias gradient of a layer whose weight gradient went through the GEMM).  grid: ceil(N / 32); block (32, 8)
__global__ void colsum_rows_kernel(const float* __restrict__ a, int M, int N, float* __restrict__ out) {
  pdl_sync();
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float v = 0.f;
  if (n < N)
    for (int m = threadIdx.y; m < M; m += 8) v += a[(size_t)m * N + n];
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    atomicAdd(&out[n], t);
  }
}

}  // namespace dsact
