// Convolution kernels of the CNN approximators (BASELINE config 5; reference networks/cnn.py:30-53: Conv2d + ReLU per
// layer, no padding, square kernels, NCHW fp32).  First CUDA path for this configuration: direct convolutions in fp32
// (bit-comparable with the reference's fp32 arithmetic up to summation order); the dense heads behind the encoder
// run through the grouped GEMM kernels of gemm_simt.cuh.
//
//   conv_fwd_kernel    y = relu(conv(x, w) + b)            one thread per output element, weights of the block's
//                                                           output channel staged in shared memory
//   conv_dgrad_kernel  dx = convT(dy, w) (.) [x > 0]       one thread per input element
//   conv_wgrad_kernel  dw = corr(x, dy), db = sum(dy)      one block per (co, ci) filter plane, reduction over
//                                                           batch and output positions
// dy is the gradient w.r.t. the layer's ReLU OUTPUT masked by the caller's chain: conv_dgrad applies the mask of the
// layer BELOW (its input x is that layer's ReLU output; x > 0 <=> the unit was active).
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace dsact {

struct ConvShape {
  int B, Cin, Hin, Win, Cout, K, S, Hout, Wout;
};

// grid: (ceil(Hout*Wout / 128), Cout, B); block 128
__global__ void conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                float* __restrict__ y, const ConvShape s) {
  pdl_sync();
  extern __shared__ float wsm[];   // [Cin][K][K] of this block's output channel
  const int co = blockIdx.y, b = blockIdx.z;
  const int nw = s.Cin * s.K * s.K;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) wsm[i] = w[(size_t)co * nw + i];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= s.Hout * s.Wout) return;
  const int oy = p / s.Wout, ox = p - oy * s.Wout;
  const float* xb = x + (size_t)b * s.Cin * s.Hin * s.Win;
  float acc = bias[co];
  for (int ci = 0; ci < s.Cin; ++ci) {
    const float* xc = xb + (size_t)ci * s.Hin * s.Win + (size_t)(oy * s.S) * s.Win + ox * s.S;
    const float* wc = wsm + ci * s.K * s.K;
    for (int ky = 0; ky < s.K; ++ky)
      for (int kx = 0; kx < s.K; ++kx) acc = fmaf(xc[ky * s.Win + kx], wc[ky * s.K + kx], acc);
  }
  y[((size_t)b * s.Cout + co) * s.Hout * s.Wout + p] = fmaxf(acc, 0.f);
}

// The same with 8 output channels per thread: the input patch is read once for eight accumulators (the one-channel form
// is bound by its loads).  grid: (ceil(Hout*Wout / 128), Cout / 8, B); block 128; dynamic smem 8 * Cin * K * K floats.
__global__ void conv_fwd8_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                 float* __restrict__ y, const ConvShape s) {
  pdl_sync();
  extern __shared__ float wsm[];   // [Cin*K*K][8]: the eight weights of one tap are adjacent (two 16-byte broadcasts)
  const int co0 = blockIdx.y * 8, b = blockIdx.z;
  const int nw = s.Cin * s.K * s.K;
  for (int i = threadIdx.x; i < 8 * nw; i += blockDim.x) {
    const int c = i / nw, tap = i - c * nw;
    wsm[tap * 8 + c] = w[(size_t)(co0 + c) * nw + tap];
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= s.Hout * s.Wout) return;
  const int oy = p / s.Wout, ox = p - oy * s.Wout;
  const float* xb = x + (size_t)b * s.Cin * s.Hin * s.Win;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = bias[co0 + c];
  for (int ci = 0; ci < s.Cin; ++ci) {
    const float* xc = xb + (size_t)ci * s.Hin * s.Win + (size_t)(oy * s.S) * s.Win + ox * s.S;
    const float* wc = wsm + (size_t)ci * s.K * s.K * 8;
    for (int ky = 0; ky < s.K; ++ky)
      for (int kx = 0; kx < s.K; ++kx) {
        const float v = xc[ky * s.Win + kx];
        const float4 w0 = *reinterpret_cast<const float4*>(wc + (ky * s.K + kx) * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wc + (ky * s.K + kx) * 8 + 4);
        acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
        acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
      }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) y[((size_t)b * s.Cout + co0 + c) * s.Hout * s.Wout + p] = fmaxf(acc[c], 0.f);
}

// dL/dx (pre-mask) then masked by x > 0 when `mask_by_x` (x is the ReLU output of the layer below; the first layer's
// input is the image: no mask, and its dx is not needed at all).  grid: (ceil(Hin*Win / 128), Cin, B); block 128
__global__ void conv_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, const float* __restrict__ x,
                                  float* __restrict__ dx, const ConvShape s, int mask_by_x) {
  pdl_sync();
  const int ci = blockIdx.y, b = blockIdx.z;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= s.Hin * s.Win) return;
  const int iy = p / s.Win, ix = p - iy * s.Win;
  const size_t xi = ((size_t)b * s.Cin + ci) * s.Hin * s.Win + p;
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
      const float* dyp = dy + (size_t)b * s.Cout * s.Hout * s.Wout + (size_t)oy * s.Wout + ox;
      const float* wp = w + ((size_t)ci * s.K + ky) * s.K + kx;
      for (int co = 0; co < s.Cout; ++co)
        acc = fmaf(dyp[(size_t)co * s.Hout * s.Wout], wp[(size_t)co * s.Cin * s.K * s.K], acc);
    }
  }
  dx[xi] = acc;
}

// dw[co][ci][ky][kx] (+)= sum_{b,oy,ox} dy[b][co][oy][ox] * x[b][ci][oy*S+ky][ox*S+kx];  db[co] (+)= sum dy (ci == 0 blocks).
// grid: (Cin, Cout); block 256; dynamic smem: K*K*8 floats of per-warp partials.  `accumulate` = += (the caller zeroed).
__global__ void conv_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                                  float* __restrict__ db, const ConvShape s) {
  pdl_sync();
  const int ci = blockIdx.x, co = blockIdx.y;
  const int KK = s.K * s.K;   // <= 16
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  float bsum = 0.f;
  const int npos = s.Hout * s.Wout;
  const long long total = (long long)s.B * npos;
  for (long long t = threadIdx.x; t < total; t += blockDim.x) {
    const int b = (int)(t / npos), p = (int)(t - (long long)b * npos);
    const int oy = p / s.Wout, ox = p - oy * s.Wout;
    const float g = dy[((size_t)b * s.Cout + co) * npos + p];
    if (g == 0.f) continue;   // ReLU-masked gradients are mostly zero
    const float* xp = x + ((size_t)b * s.Cin + ci) * s.Hin * s.Win + (size_t)(oy * s.S) * s.Win + ox * s.S;
    for (int ky = 0; ky < s.K; ++ky)
      for (int kx = 0; kx < s.K; ++kx) acc[ky * s.K + kx] = fmaf(g, xp[ky * s.Win + kx], acc[ky * s.K + kx]);
    bsum += g;
  }
  __shared__ float red[17 * 8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) red[i * 8 + wid] = v;
  }
  {
    const float v = warp_sum(bsum);
    if (lane == 0) red[16 * 8 + wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 17) {
    float v = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) v += red[threadIdx.x * 8 + k];
    if (threadIdx.x < KK) dw[((size_t)co * s.Cin + ci) * KK + threadIdx.x] = v;
    else if (threadIdx.x == 16 && ci == 0) db[co] = v;
  }
}

}  // namespace dsact
