// Persistent, warp-specialised tcgen05 / TMA tile engine for sm_100a.
//
// One kernel skeleton computes accumulator tiles  ACC[128*kCG x 256] = A[rows, K] * B[cols, K]^T
// (bf16 operands, K-major, fp32 accumulate in TMEM) and hands every tile to an epilogue policy:
//
//   EpiStats   online (max, sum-exp) row statistics of the masked score matrix  — S is never stored
//   EpiPStore  recomputes score tiles and writes the bf16 dS panel  P = incl * (a e^{S-rq} + b e^{S-rk})
//   EpiStore   plain GEMM epilogue  C = alpha * (ACC - gamma * SUB)
//
// Roles (384 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer (leader CTA only when
// kCG == 2), warp 2 = TMEM allocator, warps 4..11 = epilogue (warp%4 selects the 32-lane TMEM
// quarter, (warp-4)/4 the 128-column half of the tile).  Pipelines: smem ring full/empty (TMA <-> MMA)
// and a 2-deep TMEM accumulator ring tmem_full/tmem_empty (MMA <-> epilogue), so the epilogue of
// tile t overlaps the MMAs of tile t+1.  With kCG == 2 a CTA pair (cluster 2x1) shares the B operand:
// each CTA loads its own 128 A rows and one 128-row half of the 256-row B tile, the leader issues
// cta_group::2 MMAs with M = 256 and commits are multicast to both CTAs.
#pragma once
#include "ptx.cuh"

namespace mi {

constexpr int BLOCK_M = 128;   // accumulator rows per CTA (TMEM lanes)
constexpr int TILE_N = 256;    // accumulator columns per tile
constexpr int BLOCK_K = 64;    // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiWarps = 8;
constexpr int kTmemCols = 512;

template <int kCG>
struct Cfg {
  static constexpr int kStages = (kCG == 1) ? 4 : 6;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;          // 16 KB
  static constexpr int kBRows = TILE_N / kCG;                    // B rows this CTA loads
  static constexpr int kBBytes = kBRows * BLOCK_K * 2;           // 32 KB / 16 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;  // + slack for 1024 B alignment
  template <class Epi> static constexpr int smem_bytes() { return kSmemBytes + Epi::kEpiSmemBytes; }
};

// Work decomposition.  A "unit" is (m block, N-range split, K split); a CTA (pair) walks units
// u = pair, pair + n_pairs, ... and inside a unit walks its N tiles, each over its K blocks.
struct Sched {
  int n_mblk;     // M blocks of 128*kCG rows
  int n_ntile;    // N tiles of 256 columns
  int n_split;    // contiguous N-range splits per M block
  int n_ksplit;   // K splits (split-K GEMM), 1 otherwise
  int order;      // 0: m fastest (concurrent units share the N range), 1: ksplit, split fastest
  int k_blocks;   // number of 64-wide K blocks per tile (all segments)
  // K segments: block kb belongs to segment kb / seg_len and reads A at K block a_seg[seg] + kb % seg_len,
  // B at b_seg[seg] + kb % seg_len.  One segment = plain GEMM; more = sums of products of hi/lo splits,
  // e.g. [P_hi | P_lo] x [V ; V] or [T_hi | T_lo] x [Y ; Y] (fp32-accumulate "strict" mode).
  int seg_len;
  int a_seg[4];
  int b_seg[4];
  int a_moff[4];  // MN-major A only: M-coordinate offset of the segment (e.g. the lo half of a [hi | lo] panel)
};

struct Unit { int m, s, ks, nt0, nt1, kb0, kb1; };

__device__ __forceinline__ int num_units(const Sched& sc) { return sc.n_mblk * sc.n_split * sc.n_ksplit; }

__device__ __forceinline__ Unit decode_unit(const Sched& sc, int u) {
  Unit r;
  if (sc.order == 0) {
    r.m = u % sc.n_mblk; int t = u / sc.n_mblk;
    r.s = t % sc.n_split; r.ks = t / sc.n_split;
  } else {
    r.ks = u % sc.n_ksplit; int t = u / sc.n_ksplit;
    r.s = t % sc.n_split; r.m = t / sc.n_split;
  }
  r.nt0 = (int)(((long long)r.s * sc.n_ntile) / sc.n_split);
  r.nt1 = (int)(((long long)(r.s + 1) * sc.n_ntile) / sc.n_split);
  r.kb0 = (int)(((long long)r.ks * sc.k_blocks) / sc.n_ksplit);
  r.kb1 = (int)(((long long)(r.ks + 1) * sc.k_blocks) / sc.n_ksplit);
  return r;
}

// kAMN: the A operand is stored M-contiguous ("MN-major": A[m,k] at base[k*ld + m], i.e. the transpose
// of a row-major [K, M] matrix).  Its 128 x 64 tile is fetched as two 64(K) x 64(M) boxes and described
// to the MMA with the MN-major SWIZZLE_128B canonical layout (LBO = 8 KB between the 64-wide M halves,
// SBO = 1 KB between 8-row K groups).
template <int kCG, class Epi, bool kAMN>
__global__ void __launch_bounds__(kNumThreads, 1)
tile_engine_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const Sched sc, const typename Epi::Params ep) {
  using C = Cfg<kCG>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[C::kStages];
  __shared__ __align__(8) uint64_t empty_bar[C::kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCG == 2) ? ptx::cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);
  const int pair_id = blockIdx.x / kCG;
  const int n_pairs = gridDim.x / kCG;
  const int n_units = num_units(sc);

  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (tiles_addr - raw_addr);
  uint8_t* epi_smem = tiles + C::kStages * C::kStageBytes;     // Epi::kEpiSmemBytes of staging, if any

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full_bar[a], 1); ptx::mbar_init(&tmem_empty_bar[a], kNumEpiWarps * kCG); }
    ptx::fence_barrier_init();
  }
  __syncwarp();
  if (warp == 2) {
    ptx::tmem_alloc<kCG>(&tmem_base_slot, kTmemCols);
    ptx::tmem_relinquish<kCG>();
  }
  ptx::tc_fence_before();
  if constexpr (kCG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // Whole warps take a role branch; inside, one lane does the work and the warp reconverges at the
  // __syncwarp() so that the .aligned teardown barrier is reached convergently.
  if (warp == 0) {
   if (lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0; uint32_t phase = 0;
    const uint32_t full0_cluster = (kCG == 2) ? ptx::mapa(ptx::smem_u32(&full_bar[0]), 0) : 0u;
    for (int u = pair_id; u < n_units; u += n_pairs) {
      const Unit un = decode_unit(sc, u);
      const int a_row = (un.m * kCG + (int)cta_rank) * BLOCK_M;
      for (int nt = un.nt0; nt < un.nt1; ++nt) {
        const int b_row = nt * TILE_N + (int)cta_rank * C::kBRows;
        int seg = un.kb0 / sc.seg_len, w = un.kb0 % sc.seg_len;
        for (int kb = un.kb0; kb < un.kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1u, 1);
          uint8_t* sa = tiles + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          const int a_k = (sc.a_seg[seg] + w) * BLOCK_K;
          const int b_k = (sc.b_seg[seg] + w) * BLOCK_K;
          const int a_m = a_row + sc.a_moff[seg];
          if (++w == sc.seg_len) { w = 0; ++seg; }
          if constexpr (kCG == 1) {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
            if constexpr (kAMN) {
              ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], a_m, a_k);
              ptx::tma_load_2d(sa + C::kABytes / 2, &tmap_a, &full_bar[stage], a_m + 64, a_k);
            } else {
              ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], a_k, a_row);
            }
            ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], b_k, b_row);
          } else {
            if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
            const uint32_t bar = full0_cluster + (uint32_t)stage * 8u;
            if constexpr (kAMN) {
              ptx::tma_load_2d_2sm(sa, &tmap_a, bar, a_m, a_k);
              ptx::tma_load_2d_2sm(sa + C::kABytes / 2, &tmap_a, bar, a_m + 64, a_k);
            } else {
              ptx::tma_load_2d_2sm(sa, &tmap_a, bar, a_k, a_row);
            }
            ptx::tma_load_2d_2sm(sb, &tmap_b, bar, b_k, b_row);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
   }
   __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BLOCK_M * kCG, TILE_N) | (kAMN ? (1u << 15) : 0u);
      int stage = 0; uint32_t phase = 0; uint32_t tile_cnt = 0;
      for (int u = pair_id; u < n_units; u += n_pairs) {
        const Unit un = decode_unit(sc, u);
        for (int nt = un.nt0; nt < un.nt1; ++nt) {
          const uint32_t as = tile_cnt & 1u, aphase = (tile_cnt >> 1) & 1u;
          ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1u, 2);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * TILE_N;
          for (int kb = un.kb0; kb < un.kb1; ++kb) {
            ptx::mbar_wait(&full_bar[stage], phase, 3);
            ptx::tc_fence_after();
            const uint32_t sa = tiles_addr + stage * C::kStageBytes;
            const uint64_t da = kAMN ? ptx::make_smem_desc_mn128(sa, C::kABytes / 2) : ptx::make_smem_desc_k128(sa);
            const uint64_t db = ptx::make_smem_desc_k128(sa + C::kABytes);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // K-major: advance 16 bf16 = 32 B inside the 128 B swizzle row (+2 in 16 B units);
              // MN-major: advance 16 K rows = two 1 KB swizzle atoms (+128 in 16 B units)
              const uint64_t a_adv = kAMN ? (uint64_t)(128 * k) : (uint64_t)(2 * k);
              ptx::umma_bf16<kCG>(d_tmem, da + a_adv, db + 2 * k, idesc, (kb > un.kb0 || k > 0) ? 1u : 0u);
            }
            ptx::umma_commit<kCG>(&empty_bar[stage], 0x3);     // smem slot reusable once these MMAs retire
            if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
          }
          ptx::umma_commit<kCG>(&tmem_full_bar[as], 0x3);      // accumulator ready for the epilogue
          ++tile_cnt;
        }
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------ epilogue (8 warps)
    const uint32_t quarter = warp & 3u;
    const uint32_t half = (warp - kEpiWarp0) >> 2;
    uint32_t tile_cnt = 0;
    typename Epi::State st;
    st.stage_smem = epi_smem + (warp - kEpiWarp0) * (Epi::kEpiSmemBytes / kNumEpiWarps);
    for (int u = pair_id; u < n_units; u += n_pairs) {
      const Unit un = decode_unit(sc, u);
      const int row = (un.m * kCG + (int)cta_rank) * BLOCK_M + (int)(quarter * 32u + lane);
      Epi::unit_begin(ep, st, un, row, (int)half);
      for (int nt = un.nt0; nt < un.nt1; ++nt) {
        const uint32_t as = tile_cnt & 1u, aphase = (tile_cnt >> 1) & 1u;
        ptx::mbar_wait(&tmem_full_bar[as], aphase, 4);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const uint32_t col_in_tile = half * 128u + (uint32_t)c * 32u;
          const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + as * TILE_N + col_in_tile;
          uint32_t v[32];
          ptx::tmem_ld_32x32(taddr, v);
          ptx::tmem_ld_wait();
          Epi::chunk(ep, st, un, row, nt * TILE_N + (int)col_in_tile, c, v);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kCG == 1) ptx::mbar_arrive(&tmem_empty_bar[as]);
          else ptx::mbar_arrive_remote(&tmem_empty_bar[as], 0);
        }
        ++tile_cnt;
      }
      Epi::unit_end(ep, st, un, row, (int)half);
    }
  }

  // ---------------------------------------------------------------- teardown
  ptx::tc_fence_before();
  if constexpr (kCG == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<kCG>(tmem_base, kTmemCols);
}

// =====================================================================================
// Epilogue policies
// =====================================================================================
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

// ---- score statistics: per row online (max, sum-exp) over the negatives, negatives count, diagonal
struct EpiStats {
  static constexpr int kEpiSmemBytes = 0;
  struct Params {
    const int* sid_q;      // [q_rows] study index of accumulator rows
    const int* sid_k;      // [n_ntile*256] (padded) study index of accumulator columns
    int q_rows, k_cols;
    long long q_offset;    // column index of row 0's own sample (diagonal = q_offset + row)
    float scale;           // S = scale * acc
    float4* part;          // [n_split][2][rows_padded]  {max (log2 units), sum, count, diag}
    int rows_padded;
  };
  struct State { uint8_t* stage_smem; float m, s, cnt, diag; int sidq; };

  static __device__ __forceinline__ void unit_begin(const Params& p, State& st, const Unit&, int row, int) {
    st.m = neg_inf(); st.s = 0.f; st.cnt = 0.f; st.diag = 0.f;
    st.sidq = (row < p.q_rows) ? __ldg(p.sid_q + row) : -1;
  }
  static __device__ __forceinline__ void chunk(const Params& p, State& st, const Unit&, int row, int col0, int,
                                               const uint32_t (&v)[32]) {
    const float c2 = p.scale * kLog2e;
    const int limit = p.k_cols - col0;
    const long long dcol = p.q_offset + row - col0;
    const int4* sk4 = reinterpret_cast<const int4*>(p.sid_k + col0);
    float x[32];
    float cm = neg_inf();
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int4 sk = __ldg(sk4 + g);
      const int sks[4] = {sk.x, sk.y, sk.z, sk.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = g * 4 + j;
        const bool neg = (sks[j] != st.sidq) && (c < limit);
        x[c] = neg ? __uint_as_float(v[c]) * c2 : neg_inf();
        st.cnt += neg ? 1.f : 0.f;
        cm = fmaxf(cm, x[c]);
      }
    }
    if (dcol >= 0 && dcol < 32) {
#pragma unroll
      for (int c = 0; c < 32; ++c) if (c == (int)dcol) st.diag = __uint_as_float(v[c]) * p.scale;
    }
    const float m_new = fmaxf(st.m, cm);
    if (m_new > neg_inf()) {
      float s = st.s * ptx::ex2(st.m - m_new);
#pragma unroll
      for (int c = 0; c < 32; ++c) s += ptx::ex2(x[c] - m_new);
      st.s = s; st.m = m_new;
    }
  }
  static __device__ __forceinline__ void unit_end(const Params& p, State& st, const Unit& un, int row, int half) {
    p.part[((size_t)un.s * 2 + half) * p.rows_padded + row] = make_float4(st.m, st.s, st.cnt, st.diag);
  }
};

// ---- dS panel: P[row, col] = incl * ( wq e^{S - refq[row]} + wk e^{S - refk[col]} ), bf16 (hi [+ lo])
// Each epilogue warp stages 32 rows x 64 columns (4 KB, XOR-swizzled 16 B chunks) in shared memory and
// writes them out as full 128 B lines (4 rows per warp store) instead of 16 B per row.
struct EpiPStore {
  static constexpr int kEpiSmemBytes = kNumEpiWarps * 4096;
  struct Params {
    const int* sid_q;
    const int* sid_k;       // padded to n_ntile*256
    int q_rows, k_cols;
    long long q_offset;
    float scale;
    const float* refq;      // [q_rows] natural-log reference per row (used when use_q)
    float ln_wq;            // ln(weight) of the row term
    int use_q;
    const float* refk2;     // [n_ntile*256] (padded) column reference, log2 units, weight folded in
    int use_k;
    int include_diag;       // 1: the positive pair is part of the softmax (InfoNCE), 0: negatives only (DV)
    __nv_bfloat16* P;       // [q_rows, pitch]
    __nv_bfloat16* P_lo;    // residual panel (strict mode) or nullptr
    long long pitch;
  };
  struct State { uint8_t* stage_smem; float rq2; int sidq; };

  static __device__ __forceinline__ void unit_begin(const Params& p, State& st, const Unit&, int row, int) {
    const bool ok = row < p.q_rows;
    st.sidq = ok ? __ldg(p.sid_q + row) : -1;
    const float r = (p.use_q && ok) ? __ldg(p.refq + row) : 0.f;
    st.rq2 = (r - p.ln_wq) * kLog2e;
  }
  // write this thread's 32 packed bf16 (16 words) into its staging row, chunks [4*hc, 4*hc+4)
  static __device__ __forceinline__ void stage_row(uint8_t* smem, int lane, int hc, const uint32_t (&w)[16]) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int chunk = (hc * 4 + g) ^ (lane & 7);
      *reinterpret_cast<uint4*>(smem + lane * 128 + chunk * 16) = make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
    }
  }
  // the warp copies its 32 x 64 staging tile to global: lane -> (row = 4*it + lane/8, chunk = lane%8)
  static __device__ __forceinline__ void flush(uint8_t* smem, int lane, __nv_bfloat16* dst_row0, long long pitch, int rows_valid) {
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + (lane >> 3), ch = lane & 7;
      const uint4 val = *reinterpret_cast<const uint4*>(smem + r * 128 + ((ch ^ (r & 7)) * 16));
      if (r < rows_valid) *reinterpret_cast<uint4*>(dst_row0 + (size_t)r * pitch + ch * 8) = val;
    }
    __syncwarp();
  }
  static __device__ __forceinline__ void chunk(const Params& p, State& st, const Unit&, int row, int col0, int c_idx,
                                               const uint32_t (&v)[32]) {
    const float c2 = p.scale * kLog2e;
    const int limit = p.k_cols - col0;
    const long long dcol_ll = p.q_offset + row - col0;
    const int dcol = (p.include_diag && dcol_ll >= 0 && dcol_ll < 32) ? (int)dcol_ll : -1;
    const int4* sk4 = reinterpret_cast<const int4*>(p.sid_k + col0);
    const float4* rk4 = reinterpret_cast<const float4*>(p.refk2 + col0);
    const int lane = (int)(threadIdx.x & 31);
    float pv[32];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int4 sk = __ldg(sk4 + g);
      const int sks[4] = {sk.x, sk.y, sk.z, sk.w};
      float rks[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.use_k) { const float4 rk = __ldg(rk4 + g); rks[0] = rk.x; rks[1] = rk.y; rks[2] = rk.z; rks[3] = rk.w; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = g * 4 + j;
        const float e = __uint_as_float(v[c]) * c2;
        float val = 0.f;
        if (p.use_q) val = ptx::ex2(e - st.rq2);
        if (p.use_k) val += ptx::ex2(e - rks[j]);
        const bool incl = (c < limit) && ((sks[j] != st.sidq) || (c == dcol));
        pv[c] = incl ? val : 0.f;
      }
    }
    uint32_t hi[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) hi[c] = ptx::pack_bf16(pv[2 * c], pv[2 * c + 1]);
    const int hc = c_idx & 1;                       // which 32-column half of the 64-column staging tile
    const int row0 = row - lane;                    // first row of this warp
    const int rows_valid = p.q_rows - row0;
    const int colbase = col0 - hc * 32;
    if (p.P_lo == nullptr) {
      stage_row(st.stage_smem, lane, hc, hi);
      if (hc == 1) flush(st.stage_smem, lane, p.P + (size_t)row0 * p.pitch + colbase, p.pitch, rows_valid);
    } else {
      // strict mode: hi and lo panels go out one 32-column half at a time through the same staging tile
      uint32_t lo[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float h0 = __uint_as_float(hi[c] << 16), h1 = __uint_as_float(hi[c] & 0xffff0000u);
        lo[c] = ptx::pack_bf16(pv[2 * c] - h0, pv[2 * c + 1] - h1);
      }
      stage_row(st.stage_smem, lane, 0, hi);
      stage_row(st.stage_smem, lane, 1, lo);
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3), ch = lane & 7;
        const uint4 val = *reinterpret_cast<const uint4*>(st.stage_smem + r * 128 + ((ch ^ (r & 7)) * 16));
        __nv_bfloat16* base = (ch < 4 ? p.P : p.P_lo) + (size_t)(row0 + r) * p.pitch + col0 + (ch & 3) * 8;
        if (r < rows_valid) *reinterpret_cast<uint4*>(base) = val;
      }
      __syncwarp();
    }
  }
  static __device__ __forceinline__ void unit_end(const Params&, State&, const Unit&, int, int) {}
};

// ---- plain GEMM epilogue: C = alpha * (ACC - gamma * SUB) [+ C_prev]
struct EpiStore {
  static constexpr int kEpiSmemBytes = 0;
  struct Params {
    float* out_f32;            // optional
    __nv_bfloat16* out_bf16;   // optional
    __nv_bfloat16* out_bf16_lo;  // optional residual bf16(C - hi) (hi/lo split output), same pitch
    long long ld_out;          // pitch of out_f32
    long long ld_out16;        // pitch of out_bf16 / out_bf16_lo
    int rows, cols;
    float alpha, gamma;
    const __nv_bfloat16* sub;  // optional [rows, ld_sub]
    const __nv_bfloat16* sub_lo;  // optional residual of SUB (SUB = sub + sub_lo), same pitch
    long long ld_sub;
    int sub_row0, sub_rows;    // SUB row r applies to output row sub_row0 + r, r in [0, sub_rows)
    long long ksplit_stride;   // elements between split-K partial outputs (fp32 only)
    int accumulate;            // out_f32 += result (panel-by-panel accumulation)
  };
  struct State { uint8_t* stage_smem; };
  static __device__ __forceinline__ void unit_begin(const Params&, State&, const Unit&, int, int) {}
  static __device__ __forceinline__ void chunk(const Params& p, State&, const Unit& un, int row, int col0, int,
                                               const uint32_t (&v)[32]) {
    if (row >= p.rows || col0 >= p.cols) return;
    float o[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) o[c] = __uint_as_float(v[c]);
    const bool full = (col0 + 32 <= p.cols);
    const int srow = row - p.sub_row0;
    if (p.sub != nullptr && srow >= 0 && srow < p.sub_rows) {
      const __nv_bfloat16* s = p.sub + (size_t)srow * p.ld_sub + col0;
      if (full) {
        const uint4* s4 = reinterpret_cast<const uint4*>(s);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 q = __ldg(s4 + g);
          const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o[g * 8 + 2 * j]     -= p.gamma * __uint_as_float(w[j] << 16);
            o[g * 8 + 2 * j + 1] -= p.gamma * __uint_as_float(w[j] & 0xffff0000u);
          }
        }
      } else {
        for (int c = 0; c < 32; ++c) if (col0 + c < p.cols) o[c] -= p.gamma * __bfloat162float(s[c]);
      }
      if (p.sub_lo != nullptr) {
        const __nv_bfloat16* sl = p.sub_lo + (size_t)srow * p.ld_sub + col0;
        for (int c = 0; c < 32; ++c) if (col0 + c < p.cols) o[c] -= p.gamma * __bfloat162float(sl[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) o[c] *= p.alpha;
    if (p.out_f32 != nullptr) {
      float* d = p.out_f32 + (size_t)un.ks * p.ksplit_stride + (size_t)row * p.ld_out + col0;
      if (full) {
        float4* d4 = reinterpret_cast<float4*>(d);
        if (p.accumulate) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 prev = d4[g];
            o[4 * g] += prev.x; o[4 * g + 1] += prev.y; o[4 * g + 2] += prev.z; o[4 * g + 3] += prev.w;
          }
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) d4[g] = make_float4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
      } else {
        for (int c = 0; c < 32; ++c) if (col0 + c < p.cols) { if (p.accumulate) o[c] += d[c]; d[c] = o[c]; }
      }
    }
    if (p.out_bf16 != nullptr) {
      __nv_bfloat16* d = p.out_bf16 + (size_t)row * p.ld_out16 + col0;
      uint32_t hi[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) hi[c] = ptx::pack_bf16(o[2 * c], o[2 * c + 1]);
      if (full) {
        uint4* d4 = reinterpret_cast<uint4*>(d);
#pragma unroll
        for (int g = 0; g < 4; ++g) d4[g] = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
      } else {
        for (int c = 0; c < 32; ++c) if (col0 + c < p.cols) d[c] = __float2bfloat16(o[c]);
      }
      if (p.out_bf16_lo != nullptr) {
        __nv_bfloat16* dl = p.out_bf16_lo + (size_t)row * p.ld_out16 + col0;
        uint32_t lo[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float h0 = __uint_as_float(hi[c] << 16), h1 = __uint_as_float(hi[c] & 0xffff0000u);
          lo[c] = ptx::pack_bf16(o[2 * c] - h0, o[2 * c + 1] - h1);
        }
        if (full) {
          uint4* d4 = reinterpret_cast<uint4*>(dl);
#pragma unroll
          for (int g = 0; g < 4; ++g) d4[g] = make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
        } else {
          for (int c = 0; c < 32; ++c)
            if (col0 + c < p.cols) dl[c] = __float2bfloat16(o[c] - __bfloat162float(__float2bfloat16(o[c])));
        }
      }
    }
  }
  static __device__ __forceinline__ void unit_end(const Params&, State&, const Unit&, int, int) {}
};

}  // namespace mi
