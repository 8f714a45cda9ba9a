// Fused MLP layer-chain kernel on tcgen05 (DSACT_GEMM_BF16X3 / DSACT_GEMM_BF16).
//
// One CTA carries one 128-row block of one "pass" (an MLP applied to one input) through ALL of its layers:
//   forward chain : x -> [Linear + act] x L -> Linear            (reference networks/mlp.py:15-20)
//   dgrad chain   : dOut -> [dY W_j (.) act'(z_{j-1})] x L (-> dY W_0[:, act columns] for the actor path)
// Layer 0 takes its A operand from global memory by TMA (the bf16 hi/lo images of obs / act / dOut);
// every later layer takes A from TENSOR MEMORY: the epilogue of layer j writes act(z_j) (or dz_j) as packed
// bf16 hi/lo pairs into TMEM columns [256,512) and layer j+1 issues `tcgen05.mma` with A in TMEM, so hidden
// activations never leave the SM unless the backward pass needs them (z for act', images for wgrad).
// The weight tiles of layer j+1 are prefetched by the TMA warp while the epilogue of layer j runs.
//
// TMEM map (512 columns): [0,256) fp32 accumulator, [256,384) A hi (2 bf16 per column), [384,512) A lo.
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2..17 epilogue (lane quarter = warp % 4).
#pragma once
#include "gemm_tc.cuh"

namespace dsact {

constexpr int CH_MAX_LAYERS = DSACT_MAX_HIDDEN + 1;
constexpr int CH_MAX_PASSES = 4;
constexpr int CH_ACC_COL = 0, CH_AHI_COL = 256, CH_ALO_COL = 384;

struct ChainLayer {
  CUtensorMap mapB;          // weight image; forward: K-major (box = bn rows), dgrad: MN-major (box = 64 x 64)
  int kblocks[2];            // k-blocks of 64; layer 0 may have two A segments, later layers use [0] only
  int kB0[2];                // offset of each segment along B's reduction dimension
  int K;                     // reduction length of layers >= 1 (= width of the previous layer)
  int N, bn;                 // outputs; tile width (multiple of 16, <= 256)
  int b_mn;
  int epi, act;              // EPI_BIAS_ACT | EPI_DACT | EPI_STORE
  const float* bias;
  float* Zout;               // forward: pre-activation store, ld = N (null: not needed by a backward pass)
  const float* Zin;          // dgrad: pre-activation of the layer below, ld = N
  float* colsum;             // dgrad: bias gradient (+=)
  float* C;                  // fp32 result, ld = N (head layers)
  __nv_bfloat16* img;        // bf16 hi/lo image of the result for the weight-gradient GEMM (null: not needed)
  int img_pitch;
  long long img_plane;
};

struct ChainPass {
  CUtensorMap mapA[2];       // layer-0 A operand segments (K-major images, box = 128 rows)
  int n_layers, M, tile_start;
  ChainLayer L[CH_MAX_LAYERS];
};

struct ChainGroup {
  int n, passes;
  unsigned long long* dbg;
  ChainPass p[CH_MAX_PASSES];
};

__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__host__ __device__ inline int chain_ringA_bytes(int stages, int planes) {
  const int a = stages * planes * TC_STAGE_A, t = TC_EPI_WARPS * TR_FLOATS * 4;
  return ((a > t ? a : t) + 1023) / 1024 * 1024;
}
inline int chain_smem_bytes(int stages, int planes, int stage_b) {
  return stages * planes * stage_b + chain_ringA_bytes(stages, planes) + (2 * stages + 4) * 8 + 1024;
}

template <bool PLANES2>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_chain_kernel(const __grid_constant__ ChainGroup g, int stages, int stage_b) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space (LDS/STS)
  constexpr int planes = PLANES2 ? 2 : 1;
  // [ B ring: stages x planes x stage_b ][ layer-0 A ring: stages x planes x 16 KiB, later the epilogue's transpose scratch ]
  uint8_t* ringB = smem;
  uint8_t* ringA = smem + (size_t)stages * planes * stage_b;
  const int ringA_bytes = chain_ringA_bytes(stages, planes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ringA + ringA_bytes);
  uint64_t* full = bars;               // [stages] TMA -> MMA
  uint64_t* empty = bars + stages;     // [stages] MMA -> TMA
  uint64_t* acc_full = bars + 2 * stages;
  uint64_t* a_ready = bars + 2 * stages + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TC_STAMP(0);

  int pi = 0;
#pragma unroll
  for (int i = 1; i < CH_MAX_PASSES; ++i)
    if (i < g.n && (int)blockIdx.x >= g.p[i].tile_start) pi = i;
  const ChainPass& P = g.p[pi];
  const int m0 = (blockIdx.x - P.tile_start) * TC_BM;
  const int nl = P.n_layers;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(a_ready, TC_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TC_STAMP(1);

  if (warp == 0) {
    // ===== TMA producer: runs ahead of the epilogues, bounded only by free ring slots =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nl; ++j) {
        const ChainLayer& Lj = P.L[j];
        const int nkb = Lj.kblocks[0] + Lj.kblocks[1];
        const int b_boxes = Lj.b_mn ? (Lj.bn + 63) / 64 : 1;
        const uint32_t b_bytes = Lj.b_mn ? (uint32_t)b_boxes * 8192 : (uint32_t)Lj.bn * 128;
        const uint32_t tx = planes * (b_bytes + (j == 0 ? (uint32_t)TC_STAGE_A : 0u));
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], tx);
          const int seg = kb >= Lj.kblocks[0] ? 1 : 0;
          const int kloc = (seg ? kb - Lj.kblocks[0] : kb) * TC_BK;
          const int kB = Lj.kB0[seg] + kloc;
          uint8_t* sB = ringB + (size_t)stage * planes * stage_b;
          for (int pl = 0; pl < planes; ++pl) {
            if (j == 0)  // A ring slot = B ring slot: reuse is ordered by the same empty barrier
              tma_load_3d(ringA + (size_t)(stage * planes + pl) * TC_STAGE_A, &P.mapA[seg], &full[stage], kloc, m0, pl);
            if (Lj.b_mn) {
              for (int i = 0; i < b_boxes; ++i) tma_load_3d(sB + pl * stage_b + i * 8192, &Lj.mapB, &full[stage], 64 * i, kB, pl);
            } else {
              tma_load_3d(sB + pl * stage_b, &Lj.mapB, &full[stage], kB, 0, pl);
            }
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nl; ++j) {
        const ChainLayer& Lj = P.L[j];
        const int nkb = Lj.kblocks[0] + Lj.kblocks[1];
        const uint32_t idesc = make_idesc(TC_BM, Lj.bn, 0, Lj.b_mn);
        if (j > 0) {  // A of this layer = what the previous epilogue wrote to TMEM
          mbar_wait(a_ready, (uint32_t)((j - 1) & 1));
          tc_fence_after();
        }
        uint32_t accumulate = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (j == 0 && kb == 0) TC_STAMP(2);
          const uint32_t sB = smem_u32(ringB + (size_t)stage * planes * stage_b);
          const uint32_t sA = smem_u32(ringA + (size_t)(stage * planes) * TC_STAGE_A);
          const int ksteps = j == 0 ? 4 : min(4, (Lj.K - kb * TC_BK + 15) / 16);
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t b_off = Lj.b_mn ? k * 2048 : k * 32;
            const uint64_t b_hi = make_desc(sB + b_off, Lj.b_mn ? 8192 : 16, 1024);
            const uint64_t b_lo = make_desc(sB + stage_b + b_off, Lj.b_mn ? 8192 : 16, 1024);
            if (j == 0) {
              const uint64_t a_hi = make_desc(sA + k * 32, 16, 1024);
              tc_mma(tmem_base + CH_ACC_COL, a_hi, b_hi, idesc, accumulate);
              if (planes == 2) {
                const uint64_t a_lo = make_desc(sA + TC_STAGE_A + k * 32, 16, 1024);
                tc_mma(tmem_base + CH_ACC_COL, a_hi, b_lo, idesc, 1);
                tc_mma(tmem_base + CH_ACC_COL, a_lo, b_hi, idesc, 1);
              }
            } else {
              const uint32_t kcol = (uint32_t)(kb * 32 + k * 8);   // 2 bf16 per TMEM column
              tc_mma_ts(tmem_base + CH_ACC_COL, tmem_base + CH_AHI_COL + kcol, b_hi, idesc, accumulate);
              if (planes == 2) {
                tc_mma_ts(tmem_base + CH_ACC_COL, tmem_base + CH_AHI_COL + kcol, b_lo, idesc, 1);
                tc_mma_ts(tmem_base + CH_ACC_COL, tmem_base + CH_ALO_COL + kcol, b_hi, idesc, 1);
              }
            }
            accumulate = 1;
          }
          tc_commit(&empty[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(acc_full);
        TC_STAMP(8 + 3 * j);      // MMAs of layer j issued
      }
      TC_STAMP(3);
    }
  } else {
    // ===== epilogue =====
    const int quarter = warp & 3;
    const int sub = (warp - 2) >> 2;
    float* tr = reinterpret_cast<float*>(ringA) + (warp - 2) * TR_FLOATS;   // A ring is dead once layer 0's MMAs retired
    const int mbase = m0 + quarter * 32;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int j = 0; j < nl; ++j) {
      const ChainLayer& Lj = P.L[j];
      const bool feeds_next = j + 1 < nl;
      mbar_wait(acc_full, (uint32_t)(j & 1));
      tc_fence_after();
      if (j == 0 && threadIdx.x == 64) TC_STAMP(4);
      if (threadIdx.x == 64) TC_STAMP(9 + 3 * j);   // accumulator of layer j complete (seen by the epilogue)
      EpiArgs E;
      E.epi = Lj.epi; E.act = Lj.act; E.M = P.M; E.N = Lj.N; E.ldc = Lj.N; E.ldz = Lj.N;
      E.bias = Lj.bias; E.Zout = Lj.Zout; E.Zin = Lj.Zin; E.colsum = Lj.colsum; E.C = Lj.C;
      E.img = Lj.img; E.img_pitch = Lj.img_pitch; E.img_plane = Lj.img_plane;
      // when the next layer reads this one from TMEM, every column up to the next multiple of 16 must be written
      const int nch = (Lj.bn + 15) / 16;   // bn = N rounded up to 16: every column the next layer reads gets written
      for (int ch = sub; ch < nch; ch += TC_EPI_WARPS / 4) {
        const int c0 = ch * 16;
        float v[16];
        tc_ld16(lane_addr + CH_ACC_COL + (uint32_t)c0, v);   // v[i] = acc[row = lane][c0 + i]
        uint32_t whi[8], wlo[8];
        epi_chunk<PLANES2>(v, E, c0, mbase, lane, tr, feeds_next, whi, wlo);
        if (feeds_next) {  // next layer's A operand: packed bf16 pairs along K, hi and lo planes
          tc_st8(lane_addr + CH_AHI_COL + (uint32_t)(c0 / 2), whi);
          if (planes == 2) tc_st8(lane_addr + CH_ALO_COL + (uint32_t)(c0 / 2), wlo);
        }
      }
      if (threadIdx.x == 64) TC_STAMP(10 + 3 * j);  // epilogue of layer j done (first epilogue warp)
      if (feeds_next) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
      }
    }
  }

  if (lane == 0 && warp >= 2) { if (g.dbg) atomicMax(&g.dbg[(size_t)blockIdx.x * TC_DBG_SLOTS + 5], gtime()); }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TC_STAMP(6);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace dsact
