// tcgen05 contraction engine (sm_100a):  Y[g][co][s] = sum_k W[co][k] * Xin(g,k,s) + bias[co]
//
// Same operand generators and fused epilogue as the FP32 engine in gemm_simt.cuh, but the
// contraction runs on the 5th-gen tensor cores with FP32-grade accuracy by splitting every FP32
// operand into two FP16 terms, x = hi + lo (11 + 11 significant bits: |x - hi - lo| <= 2^-22 |x|),
// and issuing three FP16 MMAs (FP32 accumulate) per k-step:  D += Ahi*Bhi + Ahi*Blo + Alo*Bhi
// (the dropped lo*lo term is <= 2^-22 relative).  Weights are pre-scaled by a power of two per
// matrix so their lo terms stay in FP16's normal range (undone exactly in the epilogue);
// activations must satisfy |x| < 65504 (conversion saturates).  Single-pass TF32 / BF16 operands
// miss the 1e-4 parity bound, and a BF16 hi/lo split is ~8x less accurate than this.
//
// Persistent, warp-specialised CTA (one per SM), tile 256 (M) x 256 (N), K chunks of 32:
//   warps 0-3   epilogue: tcgen05.ld accumulator rows (thread = output channel; warp w owns TMEM
//               lanes 32*w..), bias / addend / ReLU, per-thread GroupNorm
//               partial sums (no shuffles).  Channels-last fp32 outputs ([column][channel]) make a
//               warp's 32 channels one 128-byte store; the fp32 [C][S] layout is transposed through smem.
//   warp  4     MMA issuer: one thread issues tcgen05.mma (M=128, N=256, K=16, kind::f16), 12 per
//               chunk; accumulators (2 x 256 fp32 columns) live in TMEM; owns TMEM alloc/dealloc
//   warp  5     A loader: one thread, cp.async.bulk (TMA engine, no tensor map) of pre-packed
//               FP16 hi/lo weight tiles, completion on the stage's mbarrier
//   warps 6-13  B producers: generate the operand tile (plain load / GroupNorm+ReLU of the
//               producer layer / pairwise op / 3x3 im2col), split to FP16 hi/lo and write it to
//               shared memory in the UMMA canonical K-major (no-swizzle) core-matrix layout.
//               (Operands that already exist as FP16 hi/lo planes are fed by TMA instead: gemm_tma.cuh.)
// 3-stage smem ring (64 KB per stage), mbarrier full/empty pipeline, tcgen05.commit releases stages.
#pragma once
#include "gemm_simt.cuh"

namespace tc {

constexpr int BN = 256;            // columns per tile
constexpr int BK = 32;             // K per pipeline stage
constexpr int STAGES = 3;
constexpr int A_HALF = 128 * BK * 2;          // one 128-row subtile, hi or lo (8 KB)
constexpr int A_SUB = 2 * A_HALF;             // hi | lo
constexpr int B_HALF = BN * BK * 2;           // 16 KB
constexpr int STAGE_BYTES = 2 * A_SUB + 2 * B_HALF;   // 64 KB
constexpr int A_LBO = 16 * 128, B_LBO = 32 * 128, SBO = 128;
constexpr int NUM_THREADS = 448;   // 4 epilogue + 1 MMA + 1 loader + 8 producer warps (128 regs/thread)
constexpr int EPI_WARPS = 4, MMA_WARP = 4, LOAD_WARP = 5;
constexpr int PRODUCER_T0 = 192;   // first producer thread
constexpr int SCRATCH_BYTES = 4 * 32 * 33 * 4;   // epilogue transpose scratch, one 32x33 fp32 block per warp
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + SCRATCH_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(cnt) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, FP16 inputs, FP32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, no-swizzle canonical layout: 8x(16 B) core matrices; LBO = stride between the two K
// core matrices of one k-step, SBO = stride between 8-row groups (cute/atom/mma_traits_sm100.hpp
// make_umma_desc<Major::K>, LayoutType::INTERLEAVE).  version = 1 (Blackwell).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}
// kind::f16 instruction descriptor: D=F32 (c_format 1), A=B=F16 (format 0), both K-major, M=128, N=256
constexpr uint32_t IDESC = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// (x0, x1) -> hi = packed f16x2 (x0 in the low half-word), lo = packed f16x2 of the residuals
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  float h0, h1;
  asm("{\n\t.reg .b16 a, b;\n\t"
      "cvt.rn.satfinite.f16.f32 a, %3;\n\t"
      "cvt.rn.satfinite.f16.f32 b, %4;\n\t"
      "mov.b32 %0, {a, b};\n\t"
      "cvt.f32.f16 %1, a;\n\t"
      "cvt.f32.f16 %2, b;\n\t}"
      : "=r"(hi), "=f"(h0), "=f"(h1)
      : "f"(x0), "f"(x1));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}

enum { OUT_CS = 0,        // fp32 Y[g][co][s]   (layout of the FP32 engine; transposed through smem)
       OUT_CL = 2 };      // fp32  Y[row][y_ms] channels-last, row = g*y_gs + column

struct TcP {
  GemmP g;             // same fields as the FP32 engine (tile width is tc::BN here)
  const uint4* Wp;     // packed weights: [kchunk][m128 tile][hi|lo][kgroup 4][m8 16][8 rows][8 k] f16
  int m_tiles;         // ceil(M / 128) (packed rows beyond M are zero)
  int k_chunks;        // ceil(K / 32)
  int mt_per_cta;      // 1 or 2 (128-row subtiles per CTA tile)
  float out_scale;     // 2^-s: undoes the power-of-two pre-scaling of the packed weights
  int out_mode;        // OUT_*
  int dbg;             // profiling experiments only (mmmot_set_debug): 1 skip epilogue work, 2 skip A loads,
                       // 4 skip B generation, 8 skip MMA issue
};

template <bool RELU>
__device__ __forceinline__ void epi_fast(uint32_t (&v)[32], float scale, float bv, float& s1, float& s2) {
#pragma unroll
  for (int j = 0; j < 32; j++) {
    float x = fmaf(__uint_as_float(v[j]), scale, bv);
    if (RELU) x = fmaxf(x, 0.f);
    v[j] = __float_as_uint(x);
    s1 += x;
    s2 = fmaf(x, x, s2);
  }
}

template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const TcP P) {
  const GemmP& p = P.g;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t bar0 = base + STAGES * STAGE_BYTES;
  // barrier slots (8 B each): full[3], empty[3], tmem_full, tmem_empty ; then tmem base (4 B)
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * STAGES), tempty_bar = bar0 + 8u * (2 * STAGES + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 2));
  float* scratch = reinterpret_cast<float*>(sm + STAGES * STAGE_BYTES + 256);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int MT = P.mt_per_cta;
  const int mgroups = (P.m_tiles + MT - 1) / MT;
  const long total_tiles = (long)p.num_tiles * mgroups;
  const int KC = P.k_chunks;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full_bar(s), 9);   // 8 producer warps + the A loader's expect_tx arrive
      mbar_init(empty_bar(s), 1);  // tcgen05.commit
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < EPI_WARPS) {
    // =============================== EPILOGUE ===============================
    const int q = warp & 3;   // TMEM lane quadrant owned by this warp
    uint32_t tphase = 0;
    float* sc = scratch + warp * (32 * 33);
    for (long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int mg = (int)(t % mgroups);
      const int nt = (int)(t / mgroups);
      int g, c0, len;
      if (p.tile_tab) { int4 tt = p.tile_tab[nt]; g = tt.x; c0 = tt.y; len = tt.z; }
      else { g = nt / p.tiles_per_group; c0 = (nt - g * p.tiles_per_group) * BN; len = min(BN, p.S - c0); }
      mbar_wait(tfull_bar, tphase);
      tphase ^= 1;
      tc_fence_after();
      for (int mt = 0; mt < MT; mt++) {
        const int co_base = (mg * MT + mt) * 128 + q * 32;
        const int co = co_base + lane;
        const bool rowok = co < p.M;
        const float bv = (rowok && p.bias) ? __ldg(p.bias + co) : 0.f;
        for (int half = 0; half < 2; half++) {
        double d1 = 0.0, d2 = 0.0;
#pragma unroll 1
        for (int cc = 0; cc < 4; cc++) {
          const int col0 = half * 128 + cc * 32;
          if (col0 >= len) break;   // warp-uniform
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * 256 + col0), v);
          if (P.dbg & 1) continue;
          float s1 = 0.f, s2 = 0.f;
          bool fast = col0 + 32 <= len;
          float bva = bv;
          if (fast && p.addend) {
            // per-detection addend, channels-last addend[det][M]: points of one detection are contiguous, so
            // a 32-column chunk almost always lies inside one detection -> fold the addend into the bias
            const int da = __ldg(p.seg + c0 + col0), db = __ldg(p.seg + c0 + col0 + 31);
            if (da == db) { if (rowok) bva += __ldg(p.addend + (long)da * p.ld_add + co); }
            else fast = false;
          }
          if (fast) {
            if (p.relu) epi_fast<true>(v, P.out_scale, bva, s1, s2);
            else epi_fast<false>(v, P.out_scale, bva, s1, s2);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) {
              float x = fmaf(__uint_as_float(v[j]), P.out_scale, bv);
              const int col = col0 + j;
              if (p.addend && rowok && col < len)
                x += __ldg(p.addend + (long)__ldg(p.seg + c0 + col) * p.ld_add + co);
              if (p.relu) x = fmaxf(x, 0.f);
              v[j] = __float_as_uint(x);
              if (col < len) { s1 += x; s2 = fmaf(x, x, s2); }
            }
          }
          d1 += (double)s1; d2 += (double)s2;
          if (!p.Y) continue;
          if (P.out_mode == OUT_CS) {
            // fp32 [C][S]: transpose the 32x32 block through smem so each store is one 128-byte line
#pragma unroll
            for (int j = 0; j < 32; j++) sc[lane * 33 + j] = __uint_as_float(v[j]);
            __syncwarp();
            const int col = col0 + lane;
            if (col < len) {
              float* dst = p.Y + (long)g * p.y_gs + (long)co_base * p.y_ms + c0 + col;
              const int rmax = min(32, p.M - co_base);
#pragma unroll 8
              for (int r = 0; r < rmax; r++) dst[(long)r * p.y_ms] = sc[r * 33 + lane];
            }
            __syncwarp();
          } else if (rowok) {
            // channels-last: a warp's 32 consecutive channels of one column = one 128-byte store
            const int nvalid = min(32, len - col0);
            {
              float* dst = p.Y + ((long)g * p.y_gs + c0 + col0) * p.y_ms + co;
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (j < nvalid) dst[(long)j * p.y_ms] = __uint_as_float(v[j]);
            }
          }
        }
        // two statistics partials per tile (one per column half), reduced in fixed order afterwards
        if (p.part && rowok) p.part[((long)nt * 2 + half) * p.M + co] = make_double2(d1, d2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
    }
  } else if (warp == MMA_WARP) {
    // =============================== MMA ISSUER ===============================
    if (lane == 0) {
      uint32_t it = 0, tcount = 0;
      for (long t = blockIdx.x; t < total_tiles; t += gridDim.x, tcount++) {
        mbar_wait(tempty_bar, (tcount & 1) ^ 1);
        tc_fence_after();
        for (int kc = 0; kc < KC; kc++, it++) {
          const int s = it % STAGES;
          mbar_wait(full_bar(s), (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = base + s * STAGE_BYTES, sb = sa + 2 * A_SUB;
#pragma unroll
          for (int mt = 0; mt < 2; mt++) {
            if (mt < MT && !(P.dbg & 8)) {
#pragma unroll
              for (int ks = 0; ks < 2; ks++) {
                const uint64_t a_hi = smem_desc(sa + mt * A_SUB + ks * 2 * A_LBO, A_LBO, SBO);
                const uint64_t a_lo = smem_desc(sa + mt * A_SUB + A_HALF + ks * 2 * A_LBO, A_LBO, SBO);
                const uint64_t b_hi = smem_desc(sb + ks * 2 * B_LBO, B_LBO, SBO);
                const uint64_t b_lo = smem_desc(sb + B_HALF + ks * 2 * B_LBO, B_LBO, SBO);
                const uint32_t d = tmem_base + (uint32_t)(mt * 256);
                umma_f16(d, a_hi, b_hi, IDESC, (kc | ks) ? 1u : 0u);
                umma_f16(d, a_hi, b_lo, IDESC, 1u);
                umma_f16(d, a_lo, b_hi, IDESC, 1u);
              }
            }
          }
          umma_commit(empty_bar(s));                 // stage free once these MMAs have read it
          if (kc == KC - 1) umma_commit(tfull_bar);  // accumulators complete
        }
      }
    }
    __syncwarp();
  } else if (warp == LOAD_WARP) {
    // =============================== A LOADER ===============================
    if (lane == 0) {
      uint32_t it = 0;
      for (long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int mg = (int)(t % mgroups);
        const int mt0 = mg * MT;
        const int nmt = min(MT, P.m_tiles - mt0);
        for (int kc = 0; kc < KC; kc++, it++) {
          const int s = it % STAGES;
          mbar_wait(empty_bar(s), ((it / STAGES) & 1) ^ 1);
          const uint32_t bytes = (uint32_t)nmt * A_SUB;
          if (P.dbg & 2) { mbar_arrive(full_bar(s)); continue; }
          mbar_expect_tx(full_bar(s), bytes);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(P.Wp) + ((size_t)kc * P.m_tiles + mt0) * A_SUB;
          bulk_g2s(base + s * STAGE_BYTES, src, bytes, full_bar(s));
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== B PRODUCERS ===============================
    // The gather of chunk i+1 (global loads into registers) is issued before chunk i is converted and
    // published, so load latency overlaps the conversion work.
    const int pt = tid - PRODUCER_T0;  // 0..255
    const long my_tiles = (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const long my_chunks = my_tiles * KC;
    long gi = 0;  // next chunk to gather

    auto tile_cols = [&](long chunk, int& g, int& c0, int& len) {
      const long t = blockIdx.x + (chunk / KC) * gridDim.x;
      const int nt = (int)(t / mgroups);
      if (p.tile_tab) { int4 tt = p.tile_tab[nt]; g = tt.x; c0 = tt.y; len = tt.z; }
      else { g = nt / p.tiles_per_group; c0 = (nt - g * p.tiles_per_group) * BN; len = min(BN, p.S - c0); }
    };
    auto arrive_full = [&](int s) {
      fence_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(s));
    };

    {
      // ---- fp32 sources: one thread per tile column, 32 k per chunk ----
      const int col = pt;
      const uint32_t row_off = (uint32_t)(col >> 3) * 128u + (uint32_t)(col & 7) * 16u;
      int g = 0, aux = 0, pi = 0;
      bool colok = false;
      long coff = 0;
      auto gather = [&](float (&v)[BK]) {
        if (gi >= my_chunks) { gi++; return; }
        const int kc = (int)(gi % KC);
        if (kc == 0) {
          int c0, len;
          tile_cols(gi, g, c0, len);
          colok = col < len;
          aux = 0;
          if (MODE == XM_DIRECT || MODE == XM_NORM_RELU) {
            coff = (long)g * p.x_gs + c0 + col;
          } else if (MODE == XM_CONV3) {
            const int s = c0 + col, hw = p.H * p.W;
            const int img = s / hw, pix = s - img * hw;
            const int y = pix / p.W, x = pix - y * p.W;
            coff = (long)img * p.Cin * hw + pix;
#pragma unroll
            for (int tp = 0; tp < 9; tp++) {
              const int yy = y + tp / 3 - 1, xx = x + tp % 3 - 1;
              if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) aux |= 1 << tp;
            }
            if (!colok) aux = 0;
          } else {
            const int s = c0 + col;
            pi = s / p.m;
            aux = p.n + (s - pi * p.m);
          }
        }
        gi++;
        if (P.dbg & 4) return;
        const int k0 = kc * BK;
        if (MODE == XM_DIRECT || MODE == XM_NORM_RELU) {
#pragma unroll
          for (int e = 0; e < BK; e++) {
            const int k = k0 + e;
            float x = 0.f;
            if (colok && k < p.K) {
              x = __ldg(p.X + coff + (long)k * p.x_ks);
              if (MODE == XM_NORM_RELU)
                x = fmaxf(fmaf(x, __ldg(p.sc + (long)g * p.K + k), __ldg(p.sh + (long)g * p.K + k)), 0.f);
            }
            v[e] = x;
          }
        } else if (MODE == XM_CONV3) {
          // fp32 NCHW input (first VGG layer); K order: k = ci*9 + tap
          int ci = k0 / 9, tap = k0 - ci * 9;
          const int hw = p.H * p.W;
#pragma unroll
          for (int e = 0; e < BK; e++) {
            float x = 0.f;
            if (ci < p.Cin && ((aux >> tap) & 1)) {
              const int d = (tap / 3 - 1) * p.W + (tap % 3 - 1);
              x = __ldg(p.X + coff + (long)ci * hw + d);
            }
            v[e] = x;
            if (++tap == 9) { tap = 0; ci++; }
          }
        } else {
#pragma unroll
          for (int e = 0; e < BK; e++) {
            const int k = k0 + e;
            float x = 0.f;
            if (colok && k < p.K) {
              const float* fr = p.X + ((long)g * p.K + k) * p.Lf;
              const float a = __ldg(fr + pi), b = __ldg(fr + aux);
              if (MODE == XM_PAIR_MUL) x = a * b;
              else if (MODE == XM_PAIR_ABS) x = fabsf((a - b) * 0.5f);
              else x = (a - b) * 0.5f;
            }
            v[e] = x;
          }
        }
      };
      // split to f16 hi / lo and publish chunk `it` in the canonical layout
      auto publish = [&](long it, const float (&v)[BK]) {
        const int s = (int)(it % STAGES);
        mbar_wait(empty_bar(s), (uint32_t)((it / STAGES) & 1) ^ 1u);
        uint8_t* bh = sm + s * STAGE_BYTES + 2 * A_SUB;
        if (!(P.dbg & 4)) {
#pragma unroll
          for (int kg = 0; kg < BK / 8; kg++) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int q = 0; q < 4; q++) split_f16x2(v[kg * 8 + 2 * q], v[kg * 8 + 2 * q + 1], h[q], l[q]);
            *reinterpret_cast<uint4*>(bh + kg * B_LBO + row_off) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(bh + B_HALF + kg * B_LBO + row_off) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
        arrive_full(s);
      };
      float va[BK], vb[BK];
      gather(va);
      for (long it = 0; it < my_chunks; it += 2) {
        gather(vb);
        publish(it, va);
        if (it + 1 < my_chunks) {
          gather(va);
          publish(it + 1, vb);
        }
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace tc

// Host launcher.  `Wp` = weights packed by mmmot_b200/weights.py::pack_tc; out_scale = 2^-s of that packing.
// GroupNorm partials: TWO per column tile (p.part[(tile*2 + half)*M + co]) -> stats_reduce(..., mult = 2).
template <int MODE>
static int gemm_tc_launch(const GemmP& g, const uint4* Wp, float out_scale, cudaStream_t st,
                          int out_mode = tc::OUT_CS) {
  if (!Wp || g.num_tiles <= 0) return MMMOT_E_ARG;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    MM_CUDA(cudaGetDevice(&dev));
    MM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  static bool attr_set = false;
  if (!attr_set) {
    MM_CUDA(cudaFuncSetAttribute(tc::gemm_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)tc::SMEM_BYTES));
    attr_set = true;
  }
  tc::TcP P;
  P.g = g;
  P.Wp = Wp;
  P.m_tiles = (g.M + 127) / 128;
  P.k_chunks = (g.K + tc::BK - 1) / tc::BK;
  P.mt_per_cta = P.m_tiles >= 2 ? 2 : 1;
  P.out_scale = out_scale;
  P.out_mode = out_mode;
  P.dbg = mm_debug_flags();
  const long mgroups = (P.m_tiles + P.mt_per_cta - 1) / P.mt_per_cta;
  const long total = (long)g.num_tiles * mgroups;
  const int grid = (int)(total < sms ? total : sms);
  tc::gemm_tc_kernel<MODE><<<grid, tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(P);
  MM_LAUNCH_CHECK();
  return 0;
}
