// tcgen05 grouped GEMM for the dense layers of the DSAC-T update (DSACT_GEMM_BF16X3 / DSACT_GEMM_BF16).
//
// Operands are bf16 "images" of the fp32 tensors: plane 0 = hi = bf16(x), plane 1 = lo = bf16(x - hi).
// BF16X3 accumulates hi*hi + hi*lo + lo*hi in fp32 (relative product error ~2^-16, enough for the 1e-4
// parity gate, SURVEY.md §7 "Parity vs speed"); BF16 uses plane 0 only.
//
// One CTA computes one 128 x BN output tile (BN <= 256) of one problem of the group:
//   warp 0      : TMA producer  (cp.async.bulk.tensor.3d, 128B swizzle, mbarrier complete_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (accumulator in TMEM, fp32)
//   warps 2..17 : epilogue      (tcgen05.ld 32x32b -> shared-memory transpose -> one output column per lane, so every
//                                global access of bias/activation/derivative/fp32/bf16-image traffic is coalesced)
// The three orientations of a linear layer never need a transposed copy: the UMMA descriptors read
// K-major or MN-major shared-memory tiles as the reduction dimension requires
//   forward  y  = x W^T   : A K-major (x image),   B K-major  (W image)
//   dgrad    dx = dy W    : A K-major (dy image),  B MN-major (W image)
//   wgrad    dW = dy^T x  : A MN-major (dy image), B MN-major (x image); split over the batch, each split
//                           stores its partial tile to a workspace slab (no atomics), reduced later.
#pragma once
#include <stdio.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_simt.cuh"  // activations, EPI_* enums
#include "kernels.cuh"    // begin_step / noise bodies of the merged prologue launch

namespace dsact {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;        // bf16 elements per k-block = one 128-byte swizzle row
constexpr int TC_MAXG = 16;
constexpr int TC_STAGE_A = TC_BM * TC_BK * 2;   // 16 KiB per plane
constexpr int TC_STAGE_B = 256 * TC_BK * 2;     // 32 KiB per plane
constexpr int TC_EPI_WARPS = 16;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;

enum { EPI_PARTIAL = 4 };  // wgrad: plain store into slab `ks`

struct TcProb {
  CUtensorMap mapA[2];      // A operand; a second K segment for cat(obs, act)
  CUtensorMap mapB;
  int kblocks[2];           // k-blocks (of 64) per A segment
  int kB0[2];               // element offset of each segment along B's reduction dimension
  int M, N, bn;             // output extents; N-tile width (multiple of 16, <= 256)
  int tiles_m, tiles_n, ksplit, tile_start;
  float* C;                 // fp32 output or null
  int ldc;
  long long split_stride;   // floats between consecutive split slabs (EPI_PARTIAL)
  const float* bias;
  float* Zout;              // pre-activation store (ld = ldc)
  const float* Zin;         // pre-activation input for the derivative
  int ldz;
  float* colsum;            // bias gradient accumulation (EPI_DACT)
  __nv_bfloat16* img;       // bf16 hi/lo image of the result, or null
  int img_pitch;
  long long img_plane;      // elements between the hi and lo planes
  int epi, act;
};

struct TcGroup {
  int n;
  int passes;               // 3 = hi*hi + hi*lo + lo*hi, 1 = hi*hi
  unsigned long long* dbg;  // optional per-CTA phase timestamps (DSACT_TC_DEBUG), 8 slots per CTA
  int tile0;                // first tile of this launch (a group may be issued as several launches of bounded size)
  TcProb p[TC_MAXG];
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Build with -DDSACT_MBAR_GUARD to turn a barrier that never completes (a protocol bug in a kernel under development)
// into a trap after ~2 s instead of a hung GPU: DSACT_NVCC_FLAGS="-DDSACT_MBAR_GUARD" python -c "import __graft_entry__ ..."
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef DSACT_MBAR_GUARD
  unsigned long long t0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680) : "memory");
    if (ok) return;
    unsigned long long t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (t1 - t0 > 2000000000ull) {
      printf("mbar_wait timeout: block %d thread %d barrier smem+%u parity %u\n", (int)blockIdx.x, (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
#else
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680)
      : "memory");
#endif
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {  // arrives on `bar` when all prior MMAs of this thread finish
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), 128-byte swizzle.
//   K-major : rows of 128 B (64 bf16 of K), 8-row groups SBO = 1024 B apart; advance K by 32 B per UMMA_K=16.
//   MN-major: k-rows of 128 B (64 bf16 of M/N), 8-k groups SBO = 1024 B apart, 64-wide M/N blocks LBO apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 x bf16.
__device__ __forceinline__ uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TC_DBG_SLOTS 48
#define TC_STAMP(slot) do { if (g.dbg) g.dbg[(size_t)blockIdx.x * TC_DBG_SLOTS + (slot)] = gtime(); } while (0)
// SM-cycle stamps of one epilogue warp's first chunk of layer 1 (slots 32..39): where a chunk's time goes
#define TC_CSTAMP(slot) do { if (g.dbg && j == 1 && k == 0 && threadIdx.x == 64) g.dbg[(size_t)blockIdx.x * TC_DBG_SLOTS + (slot)] = (unsigned long long)clock64(); } while (0)

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// ---- epilogue math ---------------------------------------------------------------------------------------
// Packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot).  The chain epilogue is bound by
// instruction issue, so every multiply-add of the activation runs on pairs of neighbouring columns.
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 pk2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 bc2(float c) { return pk2(c, c); }

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpf(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Exact-erf GELU (nn.GELU(), reference networks/mlp.py:15-20) and its derivative for a pair of pre-activations:
//   Phi(z) = 0.5 erfc(-z / sqrt 2),  erfc(x) = t P(t) exp(-x^2),  t = 1 / (1 + p x)  for x = |z| / sqrt 2 >= 0
// (Abramowitz & Stegun 7.1.26, |error of erf| <= 1.5e-7; 3.0e-7 on Phi as evaluated here in fp32 — the bf16 split
// products of this mode carry 1.5e-5).  exp(-x^2) = exp(-z^2 / 2) is also the Gaussian density up to a constant, so
// one ex2 serves Phi and phi: 2 MUFU + 10 scalar + 13 packed instructions per PAIR of elements.
template <bool WANT_D>
__device__ __forceinline__ void gelu_pair(float& z0, float& z1, float& d0, float& d1) {
  const f2 z = pk2(z0, z1);
  const f2 az = pk2(fabsf(z0), fabsf(z1));
  const f2 den = fma2(az, bc2(0.23164189f), bc2(1.0f));          // 1 + p |z| / sqrt 2, p = 0.3275911
  float n0, n1;
  upk2(den, n0, n1);
  const f2 t = pk2(rcpf(n0), rcpf(n1));
  const f2 se = mul2(mul2(z, bc2(-0.72134752044448170f)), z);     // -z^2 / 2 * log2 e
  float s0, s1;
  upk2(se, s0, s1);
  const f2 e = pk2(ex2f(s0), ex2f(s1));                            // exp(-z^2 / 2)
  f2 pl = fma2(bc2(0.5f * 1.061405429f), t, bc2(0.5f * -1.453152027f));
  pl = fma2(pl, t, bc2(0.5f * 1.421413741f));
  pl = fma2(pl, t, bc2(0.5f * -0.284496736f));
  pl = fma2(pl, t, bc2(0.5f * 0.254829592f));
  const f2 hu = mul2(mul2(pl, t), e);                              // 0.5 erfc(|z| / sqrt 2) = Phi(-|z|)
  const f2 omh = fma2(hu, bc2(-1.0f), bc2(1.0f));
  float h0, h1, o0, o1;
  upk2(hu, h0, h1);
  upk2(omh, o0, o1);
  const f2 cdf = pk2(z0 < 0.f ? h0 : o0, z1 < 0.f ? h1 : o1);
  const f2 a = mul2(z, cdf);
  if (WANT_D) {
    const f2 dd = fma2(mul2(z, bc2(0.3989422804014327f)), e, cdf);   // Phi + z phi
    upk2(dd, d0, d1);
  }
  upk2(a, z0, z1);
}
// the same arithmetic on one element (per-layer kernel: its column-layout epilogue has no registers to spare for pairs)
template <bool WANT_D>
__device__ __forceinline__ void gelu_one(float& z0, float& d0) {
  const float z = z0;
  const float t = rcpf(fmaf(fabsf(z), 0.23164189f, 1.0f));
  const float e = ex2f(z * -0.72134752044448170f * z);
  float pl = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  pl = fmaf(pl, t, 0.5f * 1.421413741f);
  pl = fmaf(pl, t, 0.5f * -0.284496736f);
  pl = fmaf(pl, t, 0.5f * 0.254829592f);
  const float hu = pl * t * e;
  const float cdf = z < 0.f ? hu : fmaf(hu, -1.0f, 1.0f);
  z0 = z * cdf;
  if (WANT_D) d0 = fmaf(z * 0.3989422804014327f, e, cdf);
}

// v[i] = act(z_i) with z_i = v[i] on entry; if D != nullptr also D[i] = act'(z_i).  The dispatch is hoisted out of
// the unrolled loops (inlining the 7-way switch per element made the kernel ~600 KB of SASS and fetch bound):
// GELU (the reference's default) and ReLU get unrolled bodies, the rest a compact loop over a private smem row.
template <bool WANT_D, int NV, bool PACKED = true>
__device__ __forceinline__ void act_fwdN(float (&v)[NV], float (&d)[NV], int act, float* row) {
  if (act == ACT_GELU) {
    if (PACKED) {
#pragma unroll
      for (int i = 0; i < NV; i += 2) gelu_pair<WANT_D>(v[i], v[i + 1], d[i], d[i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) gelu_one<WANT_D>(v[i], d[i]);
    }
  } else if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (WANT_D) d[i] = v[i] > 0.f ? 1.f : 0.f;
      v[i] = fmaxf(v[i], 0.f);
    }
  } else if (act != ACT_LINEAR) {
#pragma unroll
    for (int i = 0; i < NV; ++i) row[i] = v[i];
    if (WANT_D) {
#pragma unroll 1
      for (int i = 0; i < NV; ++i) row[i] = act_bwd(row[i], act);
#pragma unroll
      for (int i = 0; i < NV; ++i) { d[i] = row[i]; row[i] = v[i]; }
    }
#pragma unroll 1
    for (int i = 0; i < NV; ++i) row[i] = act_fwd(row[i], act);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = row[i];
  } else if (WANT_D) {
#pragma unroll
    for (int i = 0; i < NV; ++i) d[i] = 1.f;
  }
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 lo_k, __nv_bfloat16 hi_k) {  // lower k in the low half
  return (uint32_t)__bfloat16_as_ushort(lo_k) | ((uint32_t)__bfloat16_as_ushort(hi_k) << 16);
}
// hi/lo split of two neighbouring K elements straight into packed words (x0 = lower k -> low half).  The packed
// convert (F2FP.BF16.PACK_AB) runs on the ALU; two scalar F2F per element would queue on the 16-lane XU next to MUFU.
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo_k, float hi_k) {
  uint32_t w;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi_k), "f"(lo_k));
  return w;
}
__device__ __forceinline__ void split_pack2(float x0, float x1, uint32_t& whi, uint32_t& wlo) {
  whi = cvt_bf16x2(x0, x1);
  const f2 r = fma2(pk2(__uint_as_float(whi << 16), __uint_as_float(whi & 0xffff0000u)), bc2(-1.0f), pk2(x0, x1));   // x - hi, both lanes
  float r0, r1;
  upk2(r, r0, r1);
  wlo = cvt_bf16x2(r0, r1);
}

// What one epilogue chunk needs.  In the tcgen05 modes `Zout`/`Zin` carry act'(z) (computed where erf is already
// at hand) instead of z, so the backward epilogue is a load and a multiply.
struct EpiArgs {
  int epi, act, M, N, ldc, ldz;
  const float* bias;
  float* Zout;
  const float* Zin;
  float* colsum;
  float* C;
  __nv_bfloat16* img;
  int img_pitch;
  long long img_plane;
};

// One 16-column chunk of a 128-row tile.  On entry v[i] = accumulator[row = mbase + lane][col = n0c + i] (TMEM "row
// layout").  Global fp32 traffic happens in "column layout" (lane & 15 = column, lane >> 4 = which 16 of the
// warp's 32 rows, registers = rows: every load/store of a warp is two contiguous 64-byte row segments) reached by
// a transpose through the warp's private 32x17 shared-memory tile; the bf16 hi/lo image is written from the row
// layout as packed 16-byte vectors.  16 columns (not 32) keep the live register set small enough for 16 epilogue
// warps: the phase is latency bound and needs the warps.  On return, if `want_rows`, whi/wlo hold the packed bf16
// pairs of the result for this lane's row.
constexpr int TR_PITCH = 17;
constexpr int TR_FLOATS = 32 * TR_PITCH;
template <bool PLANES2>
__device__ __forceinline__ void epi_chunk(float (&v)[16], const EpiArgs& E, int n0c, int mbase, int lane, float* tr,
                                          bool want_rows, uint32_t (&whi)[8], uint32_t (&wlo)[8]) {
  const int rows_ok = max(0, min(32, E.M - mbase));
  const bool global_io = E.Zout || E.Zin || E.C || E.colsum;
  if (!global_io) {
    // on-chip only (target networks) or image-only: stay in row layout, bias by broadcast loads
    if (E.epi == EPI_BIAS_ACT || E.epi == EPI_STORE) {
      if (E.bias) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += (n0c + i < E.N) ? __ldg(E.bias + n0c + i) : 0.f;
      }
      if (E.epi == EPI_BIAS_ACT) {
        float dummy[16];
        act_fwdN<false, 16, false>(v, dummy, E.act, tr + lane * TR_PITCH);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (n0c + i < E.N) ? v[i] : 0.f;
  } else {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) tr[lane * TR_PITCH + i] = v[i];
    __syncwarp();
    const int c = lane & 15, r0 = (lane >> 4) * 16;
    const int n = n0c + c;
    const bool col_ok = n < E.N;
    const int nrows = col_ok ? max(0, min(16, rows_ok - r0)) : 0;
    const bool back = want_rows || E.img;
    float a[16], d[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) a[r] = tr[(r0 + r) * TR_PITCH + c];
    __syncwarp();  // everyone has read its column before rows are reused as scratch / overwritten
    if (E.epi == EPI_STORE || E.epi == EPI_BIAS_ACT) {
      const float bias_n = (E.bias && col_ok) ? __ldg(E.bias + n) : 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) a[r] += bias_n;
      if (E.epi == EPI_BIAS_ACT) {
        if (E.Zout) {
          act_fwdN<true, 16, false>(a, d, E.act, tr + lane * TR_PITCH);
          float* zp = E.Zout + (size_t)(mbase + r0) * E.ldc + n;
#pragma unroll
          for (int r = 0; r < 16; ++r)
            if (r < nrows) zp[(size_t)r * E.ldc] = d[r];
        } else {
          act_fwdN<false, 16, false>(a, d, E.act, tr + lane * TR_PITCH);
        }
      }
    } else if (E.epi == EPI_DACT) {
      const float* zp = E.Zin + (size_t)(mbase + r0) * E.ldz + n;
#pragma unroll
      for (int r = 0; r < 16; ++r) d[r] = (r < nrows) ? __ldg(zp + (size_t)r * E.ldz) : 0.f;
      float csum = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        a[r] *= d[r];
        csum += a[r];
      }
      if (E.colsum) {  // bias gradient of this tile: the two half-warps hold the two row halves of each column
        csum += __shfl_xor_sync(0xffffffffu, csum, 16);
        if (col_ok && lane < 16) atomicAdd(E.colsum + n, csum);
      }
    }
    if (E.C) {
      float* cp = E.C + (size_t)(mbase + r0) * E.ldc + n;
#pragma unroll
      for (int r = 0; r < 16; ++r)
        if (r < nrows) cp[(size_t)r * E.ldc] = a[r];
    }
    if (back) {  // back to row layout
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 16; ++r) tr[(r0 + r) * TR_PITCH + c] = (r < nrows) ? a[r] : 0.f;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = tr[lane * TR_PITCH + i];
    }
  }
  if (want_rows || E.img) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      split_pack2(v[2 * i], v[2 * i + 1], whi[i], wlo[i]);
    }
    if (E.img && lane < rows_ok) {  // this lane's row: 16 bf16 per plane as 16-byte vectors (pitch % 8 == 0, n0c % 16 == 0)
      __nv_bfloat16* hp = E.img + (size_t)(mbase + lane) * E.img_pitch + n0c;
      const int nvec = max(0, min(2, ((E.N + 7) / 8 * 8 - n0c + 7) / 8));
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (q < nvec) {
          reinterpret_cast<uint4*>(hp)[q] = make_uint4(whi[4 * q], whi[4 * q + 1], whi[4 * q + 2], whi[4 * q + 3]);
          if (PLANES2) reinterpret_cast<uint4*>(hp + E.img_plane)[q] = make_uint4(wlo[4 * q], wlo[4 * q + 1], wlo[4 * q + 2], wlo[4 * q + 3]);
        }
    }
  }
}

// A_MN / B_MN: operand is MN-major (reduction dimension strided in global memory).
template <bool A_MN, bool B_MN, bool PLANES2>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcGroup g, int stages, int stage_b) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzle atoms
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic on the __shared__ array keeps LDS/STS
  constexpr int planes = PLANES2 ? 2 : 1;
  const int stage_bytes = planes * (TC_STAGE_A + stage_b);  // stage_b: bytes of one B plane (widest tile of the launch)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* full = bars;             // [stages]  TMA -> MMA
  uint64_t* empty = bars + stages;   // [stages]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TC_STAMP(0);

  const int tile_id = (int)blockIdx.x + g.tile0;
  int pi = 0;
#pragma unroll
  for (int i = 1; i < TC_MAXG; ++i)
    if (i < g.n && tile_id >= g.p[i].tile_start) pi = i;
  const TcProb& P = g.p[pi];
  int local = tile_id - P.tile_start;
  const int tiles_mn = P.tiles_m * P.tiles_n;
  const int ks = local / tiles_mn;
  local -= ks * tiles_mn;
  const int m0 = (local / P.tiles_n) * TC_BM, n0 = (local % P.tiles_n) * P.bn;
  const int bn = P.bn;

  // k-block range of this CTA (split only ever applies to single-segment problems)
  const int nkb = P.kblocks[0] + P.kblocks[1];
  int kb_begin = 0, kb_end = nkb;
  if (P.ksplit > 1) {
    const int per = (nkb + P.ksplit - 1) / P.ksplit;
    kb_begin = ks * per;
    kb_end = min(nkb, kb_begin + per);
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < bn) tmem_cols <<= 1;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (warp == 2 && lane < 3) {   // descriptor prefetch (kernel parameters: independent of the preceding kernel)
    const CUtensorMap* m = lane == 0 ? &P.mapA[0] : (lane == 1 ? (P.kblocks[1] > 0 ? &P.mapA[1] : nullptr) : &P.mapB);
    if (m) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above overlapped the predecessor's tail; its memory is needed from here on
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) TC_STAMP(1);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int a_boxes = A_MN ? 2 : 1;                    // 64-wide feature blocks of a 128-row tile
      const int b_boxes = B_MN ? (bn + 63) / 64 : 1;
      const uint32_t a_bytes = A_MN ? 2u * 64 * 128 : (uint32_t)TC_STAGE_A;
      const uint32_t b_bytes = B_MN ? (uint32_t)b_boxes * 64 * 128 : (uint32_t)bn * 128;
      const uint32_t tx = planes * (a_bytes + b_bytes);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], tx);
        const int seg = kb >= P.kblocks[0] ? 1 : 0;
        const int kloc = (seg ? kb - P.kblocks[0] : kb) * TC_BK;   // offset inside the A segment
        const int kB = P.kB0[seg] + kloc;                         // offset along B's reduction dimension
        uint8_t* sA = smem + (size_t)stage * stage_bytes;
        uint8_t* sB = sA + planes * TC_STAGE_A;
        for (int pl = 0; pl < planes; ++pl) {
          if (A_MN) {
            for (int i = 0; i < a_boxes; ++i)
              tma_load_3d(sA + pl * TC_STAGE_A + i * 8192, &P.mapA[seg], &full[stage], m0 + 64 * i, kloc, pl);
          } else {
            tma_load_3d(sA + pl * TC_STAGE_A, &P.mapA[seg], &full[stage], kloc, m0, pl);
          }
          if (B_MN) {
            for (int i = 0; i < b_boxes; ++i)
              tma_load_3d(sB + pl * stage_b + i * 8192, &P.mapB, &full[stage], n0 + 64 * i, kB, pl);
          } else {
            tma_load_3d(sB + pl * stage_b, &P.mapB, &full[stage], kB, n0, pl);
          }
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TC_BM, bn, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accumulate = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (kb == kb_begin) TC_STAMP(2);
        const uint32_t sA = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t sB = sA + planes * TC_STAGE_A;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          const uint32_t a_off = A_MN ? k * 2048 : k * 32;
          const uint32_t b_off = B_MN ? k * 2048 : k * 32;
          const uint64_t a_hi = make_desc(sA + a_off, A_MN ? 8192 : 16, 1024);
          const uint64_t b_hi = make_desc(sB + b_off, B_MN ? 8192 : 16, 1024);
          tc_mma(tmem_base, a_hi, b_hi, idesc, accumulate);
          accumulate = 1;
          if (planes == 2) {
            const uint64_t a_lo = make_desc(sA + TC_STAGE_A + a_off, A_MN ? 8192 : 16, 1024);
            const uint64_t b_lo = make_desc(sB + stage_b + b_off, B_MN ? 8192 : 16, 1024);
            tc_mma(tmem_base, a_hi, b_lo, idesc, 1);
            tc_mma(tmem_base, a_lo, b_hi, idesc, 1);
          }
        }
        tc_commit(&empty[stage]);  // smem slot reusable once these MMAs retire
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      tc_commit(acc_full);
      TC_STAMP(3);
    }
  } else {
    // ===== epilogue: 16 warps; warp w may touch TMEM lanes 32*(w%4)..+31; the four warps of a lane quarter
    // take every fourth 32-column chunk.  Each phase below is fully unrolled over the 32 rows a lane holds so that
    // the loads, the activation polynomials and the stores of different rows overlap (the phase is latency bound).
    const int quarter = warp & 3;
    const int sub = (warp - 2) >> 2;                       // 0..3
    const bool have_acc = kb_begin < kb_end;
    if (have_acc) {
      mbar_wait(acc_full, 0);  // all MMAs retired: accumulator complete, pipeline smem is dead and reusable
      tc_fence_after();
      if (threadIdx.x == 64) TC_STAMP(4);
    }
    float* tr = reinterpret_cast<float*>(smem) + (warp - 2) * TR_FLOATS;  // per-warp 32x17 transpose tile
    const int nch = (bn + 15) / 16;
    const int mbase = m0 + quarter * 32;
    EpiArgs E;
    E.epi = P.epi; E.act = P.act; E.M = P.M; E.N = P.N; E.ldc = P.ldc; E.ldz = P.ldz;
    E.bias = P.bias; E.Zout = P.Zout; E.Zin = P.Zin; E.colsum = P.colsum;
    E.C = P.C ? P.C + (P.epi == EPI_PARTIAL ? (size_t)ks * P.split_stride : 0) : nullptr;
    E.img = P.img; E.img_pitch = P.img_pitch; E.img_plane = P.img_plane;
    for (int ch = sub; ch < nch; ch += TC_EPI_WARPS / 4) {
      const int c0 = ch * 16;
      float v[16];
      if (have_acc) {
        tc_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      uint32_t whi[8], wlo[8];
      epi_chunk<PLANES2>(v, E, n0 + c0, mbase, lane, tr, false, whi, wlo);
    }
  }

  if (lane == 0 && warp >= 2) { if (g.dbg) atomicMax(&g.dbg[(size_t)blockIdx.x * TC_DBG_SLOTS + 5], gtime()); }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TC_STAMP(6);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// fp32 -> bf16 hi/lo image conversion for tensors that elementwise kernels (or the optimiser) produce.
// Up to two column segments let cat(obs, act)-shaped weights land with the act block on a 64-column boundary.
struct ImgJob {
  const float* src;
  __nv_bfloat16* dst;
  int rows, ld_src;
  int seg_w[2], seg_src0[2], seg_dst0[2];
  int pitch;            // image row pitch (elements); columns not covered by a segment are zero-filled up to `fill_w`
  int fill_w;
  long long plane;
  int block_start;
};
constexpr int IMG_MAXJ = 32;   // every weight of the six networks + a caller-supplied batch in one launch
struct ImgGroup {
  int n, planes;
  ImgJob j[IMG_MAXJ];
};
// One thread converts 8 consecutive columns of one row (one 16-byte store per plane).
__device__ __forceinline__ void image_body(const ImgGroup& g, int block, int img_blocks) {
  int ji = 0;
#pragma unroll
  for (int i = 1; i < IMG_MAXJ; ++i)
    if (i < g.n && block >= g.j[i].block_start) ji = i;
  const ImgJob& J = g.j[ji];
  const int vec_per_row = J.pitch >> 3;
  const int total = J.rows * vec_per_row;
  const int nblocks = (ji + 1 < g.n ? g.j[ji + 1].block_start : img_blocks) - J.block_start;
  for (int i = (block - J.block_start) * blockDim.x + threadIdx.x; i < total; i += nblocks * blockDim.x) {
    const int r = i / vec_per_row, c0 = (i - r * vec_per_row) * 8;
    const float* src = J.src + (size_t)r * J.ld_src;
    uint32_t whi[4], wlo[4];
    float x[8];
    const uintptr_t addr = reinterpret_cast<uintptr_t>(src + c0);
    if (c0 + 8 <= J.seg_w[0] && (addr & 7) == 0) {   // the common case: eight columns of the first segment, vector loads
      if ((addr & 15) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src + c0)), b = __ldg(reinterpret_cast<const float4*>(src + c0) + 1);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 a = __ldg(reinterpret_cast<const float2*>(src + c0) + k);
          x[2 * k] = a.x; x[2 * k + 1] = a.y;
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = c0 + e;
        float v = 0.f;
        if (c < J.seg_w[0]) v = __ldg(src + c);
        else if (c >= J.seg_dst0[1] && c < J.seg_dst0[1] + J.seg_w[1]) v = __ldg(src + J.seg_src0[1] + (c - J.seg_dst0[1]));
        x[e] = v;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) split_pack2(x[2 * k], x[2 * k + 1], whi[k], wlo[k]);
    __nv_bfloat16* dst = J.dst + (size_t)r * J.pitch + c0;
    *reinterpret_cast<uint4*>(dst) = make_uint4(whi[0], whi[1], whi[2], whi[3]);
    if (g.planes == 2) *reinterpret_cast<uint4*>(dst + J.plane) = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
  }
}

__global__ void image_kernel(const __grid_constant__ ImgGroup g) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  image_body(g, (int)blockIdx.x, (int)gridDim.x);
}

// Start-of-step work that depends on nothing of the step itself, as ONE launch: weight (and caller-batch) images, the
// accumulator / gradient clears (begin_step_kernel) and the device noise (noise_kernel), each on its own range of blocks.
struct PrologueArgs {
  int img_blocks, zero_blocks, noise_blocks;
  float *state, *grads;
  long long n_grads;
  float *eps1, *eps2, *z3, *z4;
  int B, A;
  unsigned long long seed;
  AdamHyper hy;   // this step's Adam scalars are formed here (last block of the clear range), see adam_scalars_stamp
};
__global__ void step_prologue_kernel(const __grid_constant__ ImgGroup g, const __grid_constant__ PrologueArgs p) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int b = (int)blockIdx.x;
  if (b < p.img_blocks) image_body(g, b, p.img_blocks);
  else if (b < p.img_blocks + p.zero_blocks) {
    if (b == p.img_blocks + p.zero_blocks - 1) adam_scalars_stamp(p.state, p.hy);
    begin_step_body(p.state, p.grads, p.n_grads, b - p.img_blocks, p.zero_blocks);
  }
  else noise_body(p.eps1, p.eps2, p.z3, p.z4, p.B, p.A, p.seed, p.state, b - p.img_blocks - p.zero_blocks, p.noise_blocks);
}

// Sum the wgrad split slabs into the flat gradient buffer (which already holds the bias gradients).
__global__ void grad_reduce_kernel(float* __restrict__ grads, const float* __restrict__ slabs, long long n, int nslabs,
                                   long long slab_stride) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const bool vec = (slab_stride & 3) == 0;
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 s = reinterpret_cast<const float4*>(grads)[i];
    for (int k = 0; k < nslabs; ++k) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(slabs + (size_t)k * slab_stride) + i);
      s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
    }
    reinterpret_cast<float4*>(grads)[i] = s;
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = grads[i];
    for (int k = 0; k < nslabs; ++k) s += slabs[(size_t)k * slab_stride + i];
    grads[i] = s;
  }
}

}  // namespace dsact
