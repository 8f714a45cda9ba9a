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
// TMEM map (512 columns): two 256-column buffers used alternately.
// Layer j accumulates into buffer j & 1; its epilogue converts the accumulator IN PLACE into the next layer's A operand
// (the 16 fp32 columns of a chunk become 8 columns of packed bf16 hi pairs + 8 columns of lo pairs: the same cells),
// and layer j + 1, accumulating into the other buffer, issues the MMAs of k-block kb as soon as the four chunks of that
// k-block have been written (one mbarrier per k-block, 16 warp arrivals): the tensor pipe trails the epilogue by one
// k-block instead of waiting for the whole layer, so MMA time disappears behind epilogue time.
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2..17 epilogue (lane quarter = warp % 4).
#pragma once
#include "gemm_tc.cuh"

namespace dsact {

constexpr int CH_MAX_LAYERS = DSACT_MAX_HIDDEN + 1;
constexpr int CH_MAX_PASSES = 4;

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
  CUtensorMap mapZ;          // fp32 [M, N] act'(z): stored by the forward epilogue (Zout) / loaded by the dgrad one (Zin); box 16 x 32
  CUtensorMap mapImg;        // bf16 [2][M][pitch] image of the result; box 16 x 32 x 1
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Per epilogue warp: [ fp32 tile 32x16 (act' out, or act' in buffer 0) 2 KiB ][ image hi 1 KiB ][ image lo 1 KiB ] inside the
// (dead after layer 0) A ring, plus a second act'-in buffer (2 KiB) behind it for the dgrad prefetch.
constexpr int CH_STAGE_WARP = 4096, CH_ZIN1_WARP = 2048;
__host__ __device__ inline int chain_ringA_bytes(int stages, int planes) {
  const int a = stages * planes * TC_STAGE_A, t = TC_EPI_WARPS * CH_STAGE_WARP;
  return ((a > t ? a : t) + 1023) / 1024 * 1024;
}
inline int chain_smem_bytes(int stages, int planes, int stage_b) {
  return stages * planes * stage_b + chain_ringA_bytes(stages, planes) + TC_EPI_WARPS * CH_ZIN1_WARP +
         (2 * stages + 8 + 2 * TC_EPI_WARPS) * 8 + 1024;
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
  uint8_t* zin1 = ringA + ringA_bytes;                                   // [warps] second act' input buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(zin1 + TC_EPI_WARPS * CH_ZIN1_WARP);
  uint64_t* full = bars;               // [stages] TMA -> MMA
  uint64_t* empty = bars + stages;     // [stages] MMA -> TMA
  uint64_t* acc_full = bars + 2 * stages;          // MMA -> epilogue: accumulator of the layer complete
  uint64_t* a_ready = bars + 2 * stages + 1;       // [4] epilogue -> MMA: the chunks of k-block kb of the next layer's A operand
                                                   //     are in tensor memory
  uint64_t* zbar = a_ready + 4;             // [warps][2] act' tile arrival
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(zbar + 2 * TC_EPI_WARPS);

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
    for (int i = 0; i < 4; ++i) mbar_init(&a_ready[i], TC_EPI_WARPS);
    for (int i = 0; i < 2 * TC_EPI_WARPS; ++i) mbar_init(&zbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (warp == 2) {   // descriptor prefetch: every tensor map this CTA will use (they are kernel parameters: no dependency on
                     // the preceding kernel), so that neither the first tile loads nor the epilogue's TMA stores wait for a fetch
    for (int i = lane; i < 2 + 3 * nl; i += 32) {
      const CUtensorMap* m = nullptr;
      if (i == 0) m = &P.mapA[0];
      else if (i == 1) { if (P.L[0].kblocks[1] > 0) m = &P.mapA[1]; }
      else {
        const ChainLayer& Lq = P.L[(i - 2) / 3];
        const int which = (i - 2) % 3;
        if (which == 0) m = &Lq.mapB;
        else if (which == 1) { if (Lq.Zout || Lq.Zin) m = &Lq.mapZ; }
        else if (Lq.img) m = &Lq.mapImg;
      }
      if (m) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");            // programmatic dependent launch, see gemm_tc.cuh
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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
        // accumulator / A-operand columns of this layer (see the TMEM map at the top)
        const uint32_t acc = tmem_base + (uint32_t)((j & 1) * 256);
        const uint32_t abuf = tmem_base + (uint32_t)(((j - 1) & 1) * 256);   // streamed: where layer j - 1 accumulated
        uint32_t accumulate = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          if (j > 0) mbar_wait(&a_ready[kb], (uint32_t)((j - 1) & 1));   // chunks 4 kb .. 4 kb + 3 of the operand
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
              tc_mma(acc, a_hi, b_hi, idesc, accumulate);
              if (planes == 2) {
                const uint64_t a_lo = make_desc(sA + TC_STAGE_A + k * 32, 16, 1024);
                tc_mma(acc, a_hi, b_lo, idesc, 1);
                tc_mma(acc, a_lo, b_hi, idesc, 1);
              }
            } else {
              // 16 K elements = one epilogue chunk = 8 TMEM columns of packed bf16 pairs per plane
              const uint32_t a_hi = abuf + (uint32_t)((kb * 4 + k) * 16), a_lo = a_hi + 8;
              tc_mma_ts(acc, a_hi, b_hi, idesc, accumulate);
              if (planes == 2) {
                tc_mma_ts(acc, a_hi, b_lo, idesc, 1);
                tc_mma_ts(acc, a_lo, b_hi, idesc, 1);
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
    // ===== epilogue: everything stays in the TMEM row layout (lane = row).  Results that the backward pass needs
    // go through a per-warp shared-memory tile and leave with TMA stores; act' tiles come in with TMA loads that are
    // issued one chunk ahead; nothing here touches the LSU global path except the tiny head outputs. =====
    const int quarter = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int ew = warp - 2;
    uint8_t* st = ringA + (size_t)ew * CH_STAGE_WARP;          // A ring is dead once layer 0's MMAs retired
    float* st_f32 = reinterpret_cast<float*>(st);              // 32 x 16 fp32
    uint32_t* st_hi = reinterpret_cast<uint32_t*>(st + 2048);  // 32 x 8 packed bf16 pairs
    uint32_t* st_lo = reinterpret_cast<uint32_t*>(st + 3072);
    // act' tiles: buffer 0 = the fp32 staging tile, buffer 1 = this warp's slice of zin1.  Kept as (base, byte offset)
    // so that the accesses stay LDS/STS (an array of pointers decays to generic loads).
    float* const zin1w = reinterpret_cast<float*>(zin1 + (size_t)ew * CH_ZIN1_WARP);
    const int zin1_off = (int)((zin1 + (size_t)ew * CH_ZIN1_WARP) - st) / 4;   // in floats, relative to st_f32
#define ZIN_BUF(i) (st_f32 + ((i) ? zin1_off : 0))
    uint64_t* zb = zbar + 2 * ew;
    uint32_t zphase[2] = {0, 0};
    // shared-memory side of the TMA tiles is swizzled (tc_host.cuh): 16-byte chunk index of this lane's row
    const int sw64 = (lane >> 1) & 3, sw32 = (lane >> 2) & 1;
    const int mbase = m0 + quarter * 32;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int j = 0; j < nl; ++j) {
      const ChainLayer& Lj = P.L[j];
      const bool feeds_next = j + 1 < nl;
      // bias of this warp's (up to four) chunks -> the warp's private row of the act' buffer (forward passes never load
      // act' tiles), while the MMAs of the layer are still running; the chunks then read it as broadcast float4s
      const bool has_bias = (Lj.epi == EPI_BIAS_ACT || Lj.epi == EPI_STORE) && Lj.bias;
      float* bw = zin1w;
      if (has_bias) {
        __syncwarp();
        for (int t = lane; t < 64; t += 32) {
          const int n = (sub + (TC_EPI_WARPS / 4) * (t >> 4)) * 16 + (t & 15);
          bw[t] = n < Lj.N ? __ldg(Lj.bias + n) : 0.f;
        }
        __syncwarp();
      }
      mbar_wait(acc_full, (uint32_t)(j & 1));
      tc_fence_after();
      const uint32_t acc_addr = lane_addr + (uint32_t)((j & 1) * 256);
      if (j == 0 && threadIdx.x == 64) TC_STAMP(4);
      if (threadIdx.x == 64) TC_STAMP(9 + 3 * j);   // accumulator of layer j complete (seen by the epilogue)
      const int epi = Lj.epi, act = Lj.act;
      const int nch = (Lj.bn + 15) / 16;   // bn = N rounded up to 16: every column the next layer reads gets written
      const bool dact = epi == EPI_DACT;
      if (dact && sub < nch) {  // act' tile of this warp's first chunk
        if (lane == 0) {
          tma_store_wait_read();  // buffer 0 doubles as the fp32 store tile
          mbar_expect_tx(&zb[0], 2048);
          tma_load_2d(ZIN_BUF(0), &Lj.mapZ, &zb[0], sub * 16, mbase);
        }
      }
      int k = 0;
      for (int ch = sub; ch < nch; ch += TC_EPI_WARPS / 4, ++k) {
        const int c0 = ch * 16;
        float v[16];
        TC_CSTAMP(32);
        tc_ld16(acc_addr + (uint32_t)c0, v);   // v[i] = acc[row = lane][c0 + i]
        TC_CSTAMP(33);
        if (dact) {
          const int nxt = ch + TC_EPI_WARPS / 4;
          __syncwarp();  // every lane is done with the buffer the prefetch overwrites
          if (nxt < nch && lane == 0) {  // prefetch the next act' tile into the other buffer
            if (((k + 1) & 1) == 0) tma_store_wait_read();
            mbar_expect_tx(&zb[(k + 1) & 1], 2048);
            tma_load_2d(ZIN_BUF((k + 1) & 1), &Lj.mapZ, &zb[(k + 1) & 1], nxt * 16, mbase);
          }
          mbar_wait(&zb[k & 1], zphase[k & 1]);
          zphase[k & 1] ^= 1;
          const float4* zr = reinterpret_cast<const float4*>(ZIN_BUF(k & 1) + lane * 16);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 d = zr[q ^ sw64];   // rows >= M and columns >= N arrive as zeros (TMA out-of-bounds fill)
            upk2(mul2(pk2(v[4 * q], v[4 * q + 1]), pk2(d.x, d.y)), v[4 * q], v[4 * q + 1]);
            upk2(mul2(pk2(v[4 * q + 2], v[4 * q + 3]), pk2(d.z, d.w)), v[4 * q + 2], v[4 * q + 3]);
          }
          if (Lj.colsum) {  // bias gradient: column sums over the warp's 32 rows by a reduce-scatter of the 16 columns
            float r[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = v[i];
            int off = 0;
#pragma unroll
            for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
              const bool up = lane & bit;
#pragma unroll
              for (int i = 0; i < w; ++i) {
                const float send = up ? r[i] : r[i + w];
                const float recv = __shfl_xor_sync(0xffffffffu, send, bit);
                r[i] = (up ? r[i + w] : r[i]) + recv;
              }
              off += up ? w : 0;
            }
            r[0] += __shfl_xor_sync(0xffffffffu, r[0], 1);
            if ((lane & 1) == 0 && c0 + off < Lj.N) atomicAdd(Lj.colsum + c0 + off, r[0]);
          }
        }
        float d[16];
        const bool st_z = epi == EPI_BIAS_ACT && Lj.Zout;
        // The previous chunk's TMA stores must have finished reading this warp's tile before it is written again.  Only
        // the generic activations use the tile as scratch; for GELU/ReLU/linear the wait moves behind the math, which
        // then overlaps the store engine's read.
        const bool scratch = epi == EPI_BIAS_ACT && act != ACT_GELU && act != ACT_RELU && act != ACT_LINEAR;
        if ((st_z || Lj.img) && scratch) {
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
        }
        if (epi == EPI_BIAS_ACT || epi == EPI_STORE) {
          if (has_bias) {
            const float4* bp = reinterpret_cast<const float4*>(bw + k * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b = bp[q];
              upk2(add2(pk2(v[4 * q], v[4 * q + 1]), pk2(b.x, b.y)), v[4 * q], v[4 * q + 1]);
              upk2(add2(pk2(v[4 * q + 2], v[4 * q + 3]), pk2(b.z, b.w)), v[4 * q + 2], v[4 * q + 3]);
            }
          }
          if (epi == EPI_BIAS_ACT) {
            if (Lj.Zout) act_fwdN<true, 16>(v, d, act, st_f32 + lane * 16);
            else act_fwdN<false, 16>(v, d, act, st_f32 + lane * 16);
          }
        }
        TC_CSTAMP(34);
        if (c0 + 16 > Lj.N) {   // partial chunk: padding columns feed the next layer as zeros
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (c0 + i < Lj.N) ? v[i] : 0.f;
        }
        if (Lj.C) {  // narrow head outputs (N = 2, 2A, A): direct stores
          if (mbase + lane < P.M) {
            float* cp = Lj.C + (size_t)(mbase + lane) * Lj.N + c0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + i < Lj.N) cp[i] = v[i];
          }
        }
        uint32_t whi[8], wlo[8];
        if (feeds_next || Lj.img) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            split_pack2(v[2 * i], v[2 * i + 1], whi[i], wlo[i]);
          }
        }
        TC_CSTAMP(35);
        if (st_z || Lj.img) {  // stage in shared memory (row layout), one lane issues the TMA stores (they clip at M and N)
          if (!scratch) {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
          }
          if (st_z) {
            float4* zr = reinterpret_cast<float4*>(st_f32 + lane * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) zr[q ^ sw64] = make_float4(d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3]);
          }
          if (Lj.img) {
            uint4* hr = reinterpret_cast<uint4*>(st_hi + lane * 8);
            hr[sw32] = make_uint4(whi[0], whi[1], whi[2], whi[3]);
            hr[sw32 ^ 1] = make_uint4(whi[4], whi[5], whi[6], whi[7]);
            if (PLANES2) {
              uint4* lr = reinterpret_cast<uint4*>(st_lo + lane * 8);
              lr[sw32] = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
              lr[sw32 ^ 1] = make_uint4(wlo[4], wlo[5], wlo[6], wlo[7]);
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (st_z) tma_store_2d(&Lj.mapZ, st_f32, c0, mbase);
            if (Lj.img) tma_store_3d(&Lj.mapImg, st_hi, c0, mbase, 0);   // box depth = planes: hi and lo in one request
            tma_store_commit();
          }
        }
        TC_CSTAMP(36);
        if (feeds_next) {  // next layer's A operand: packed bf16 pairs along K, hi and lo planes
          // in place: the chunk's 16 accumulator columns become 8 columns of hi pairs + 8 of lo pairs
          tc_st8(acc_addr + (uint32_t)c0, whi);
          if (planes == 2) tc_st8(acc_addr + (uint32_t)c0 + 8, wlo);
        }
        if (feeds_next) {   // this warp's share of k-block k of the next layer's operand is in tensor memory
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_ready[k]);
        }
        TC_CSTAMP(37);
      }
      if (g.dbg && j == 1 && threadIdx.x == 64) g.dbg[(size_t)blockIdx.x * TC_DBG_SLOTS + 38] = (unsigned long long)clock64();
      if (threadIdx.x == 64) TC_STAMP(10 + 3 * j);  // epilogue of layer j done (first epilogue warp)
      if (feeds_next) {   // k-blocks in which this warp had no chunk (narrow layers): every barrier sees every warp once
        for (; k < 4; ++k)
          if (lane == 0) mbar_arrive(&a_ready[k]);
      }
    }
    if (lane == 0) tma_store_wait_read();   // shared memory must stay valid until the last stores have read it
#undef ZIN_BUF
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
