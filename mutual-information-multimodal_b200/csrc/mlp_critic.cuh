// Part of libmi_b200.so: included by mi_b200.cu INSIDE its anonymous namespace, after the tile-engine launchers
// (Bump, GemmArgs, run_gemm, launch_engine, the statistics kernels).  Not a standalone header.
#pragma once

// ============================================================================================
// Concat-MLP critic — the reference's own mi_discriminator = make_mlp(1536, [1024, 512])
// (main_utils.py:77, model.py:18-32) on every (image i, text j) pair, without the pair tensor.
//
//   layer 1 is separable:  W1 [x_i ; y_j] + b1 = (W1x x_i + b1) + W1y y_j = A_i + C_j     (two [B, H1] GEMMs)
//   per pair (i, j):       h1 = relu(A_i + C_j);  z2 = W2 h1 + b2;  S_ij = w3 . relu(z2) + b3
//
// Pairs are processed in row panels (rows i in [r0, r0 + R), all columns j; pair index p = (i - r0) B + j):
//   forward   H[p,:] = bf16 h1  ->  tile engine  Z2 = H W2^T  with the EpiMlpFwd epilogue  ->  S (fp32 [B, B])
//   loss      row statistics of S over the negatives mask -> DV / InfoNCE exactly as for the separable critics; G = dL/dS
//   backward  recompute Z2 with the EpiMlpDz epilogue -> dZ2 panel (bf16), dw3, db2;
//             dW2 += dZ2^T H          (both operands read MN-major, contraction over the pairs, split-K)
//             dZ1  = (dZ2 W2) . [H > 0]  (W2 read MN-major; ReLU mask in the epilogue), reduced over j -> dA_i, over i -> dC_j
//             dX = dA W1x, dY = dC W1y, dW1 = [dA^T X | dC^T Y], db1 = sum_i dA_i
// Strict ("fp32-accumulate") mode keeps every bf16 operand as a hi/lo pair and adds the cross terms as K segments.
// ============================================================================================
// fp32 [R, C] (pitch ld_in) -> bf16 [R, pitch]: hi in columns [0, C), lo (split == 2) in [Cp, Cp + C), zeros elsewhere
__global__ void split_f32_kernel(const float* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out, long long pitch,
                                 int split, long long Cp, long long R, long long C) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= R * pitch) return;
  const long long r = idx / pitch, c = idx - r * pitch;
  float v = 0.f;
  if (c < C) v = __bfloat162float(__float2bfloat16(in[r * ld_in + c]));
  else if (split == 2 && c >= Cp && c < Cp + C) {
    const float x = in[r * ld_in + (c - Cp)];
    v = x - __bfloat162float(__float2bfloat16(x));
  }
  out[idx] = __float2bfloat16(v);
}
__global__ void pad_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n, long long n_pad) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n_pad) dst[i] = (i < n) ? src[i] : 0.f;
}
// H[p, k] = relu(A[r0 + p / Bk, k] + C[p % Bk, k]) as bf16 hi (+ lo at column Hp), 8 columns per thread
__global__ void mlp_gen_kernel(const float* __restrict__ A, const float* __restrict__ Cm, long long H1, long long Bk, long long r0,
                               long long P, __nv_bfloat16* __restrict__ H, long long pitch, long long Hp, int split) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long h8 = H1 >> 3;
  if (idx >= P * h8) return;
  const long long p = idx / h8, k = (idx - p * h8) << 3;
  const long long i = r0 + p / Bk, j = p % Bk;
  const float4* a4 = reinterpret_cast<const float4*>(A + i * H1 + k);
  const float4* c4 = reinterpret_cast<const float4*>(Cm + j * H1 + k);
  const float4 a0 = __ldg(a4), a1 = __ldg(a4 + 1), c0 = __ldg(c4), c1 = __ldg(c4 + 1);
  const float v[8] = {fmaxf(a0.x + c0.x, 0.f), fmaxf(a0.y + c0.y, 0.f), fmaxf(a0.z + c0.z, 0.f), fmaxf(a0.w + c0.w, 0.f),
                      fmaxf(a1.x + c1.x, 0.f), fmaxf(a1.y + c1.y, 0.f), fmaxf(a1.z + c1.z, 0.f), fmaxf(a1.w + c1.w, 0.f)};
  uint32_t hi[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) hi[t] = ptx::pack_bf16(v[2 * t], v[2 * t + 1]);
  *reinterpret_cast<uint4*>(H + p * pitch + k) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (split == 2) {
    uint32_t lo[4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
      lo[t] = ptx::pack_bf16(v[2 * t] - __uint_as_float(hi[t] << 16), v[2 * t + 1] - __uint_as_float(hi[t] & 0xffff0000u));
    *reinterpret_cast<uint4*>(H + p * pitch + Hp + k) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}
// S[r0 * Bk + p] = b3 + sum of the column-quarter partials
__global__ void mlp_logit_merge_kernel(const float* __restrict__ part, int rows_padded, long long P, const float* __restrict__ b3,
                                       float* __restrict__ S) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= P) return;
  float a = b3[0];
#pragma unroll
  for (int q = 0; q < mi::kColQuarters; ++q) a += part[(size_t)q * rows_padded + p];
  S[p] = a;
}
// one warp per row of S: {lse over the negatives, #negatives, S_ii, lse over negatives and the positive}
__global__ void mlp_row_stats_kernel(const float* __restrict__ S, const int* __restrict__ sid, long long B, float4* __restrict__ row_out) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const int my = sid[row];
  float m = mi::neg_inf(), s = 0.f, cnt = 0.f;
  for (long long j = lane; j < B; j += 32) {
    if (sid[j] != my) {                                   // main_utils.py:105 (the diagonal has equal ids)
      const float v = S[row * B + j];
      const float mn = fmaxf(m, v);
      s = s * expf(m - mn) + expf(v - mn); m = mn; cnt += 1.f;
    }
  }
  for (int o = 16; o; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    const float mn = fmaxf(m, m2);
    if (mn > mi::neg_inf()) s = s * expf(m - mn) + s2 * expf(m2 - mn);
    m = mn;
  }
  if (lane == 0) {
    const float diag = S[row * B + row];
    const float lse_neg = (cnt > 0.f && s > 0.f) ? m + logf(s) : mi::neg_inf();
    const float hi = fmaxf(lse_neg, diag), lo = fminf(lse_neg, diag);
    row_out[row] = make_float4(lse_neg, cnt, diag, hi + log1pf(expf(lo - hi)));
  }
}
// G = dL/dS.  dv / infonce (mi_critics.py:3-23): softmax over ALL negatives, -1/B on the diagonal;
// row InfoNCE: (1/B) softmax over {positive} u negatives of the row, minus 1/B on the diagonal.
__global__ void mlp_g_kernel(const float* __restrict__ S, const int* __restrict__ sid, long long B, int dv_like,
                             const float* __restrict__ lse, const float4* __restrict__ rows, float* __restrict__ G) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= B * B) return;
  const long long i = idx / B, j = idx - i * B;
  const float inv_b = 1.f / static_cast<float>(B);
  const bool neg = sid[i] != sid[j];
  float g = 0.f;
  if (dv_like) g = neg ? expf(S[idx] - lse[0]) : (i == j ? -inv_b : 0.f);
  else if (neg || i == j) g = inv_b * expf(S[idx] - rows[i].w) - (i == j ? inv_b : 0.f);
  G[idx] = g;
}
// One read of the dZ1 panel gives both reductions.  A block owns a slab [32 rows il] x [all j] x [64 columns]:
// thread = (column pair, j group); for every j it loads the 32 il values (32 independent loads in flight), adds them into
// its 32 register accumulators (-> dA[il]) and sends their sum to dC[j] with one atomicAdd per column.
template <typename T>
__global__ void __launch_bounds__(256) mlp_reduce_both_kernel(const T* __restrict__ dz1, long long ld, long long H1, long long Bk,
                                                              long long R, long long r0, float* __restrict__ dA, float* __restrict__ dC) {
  constexpr int IL = 32, JG = 8;
  __shared__ float sh[IL][64];
  const int kk = threadIdx.x & 31, jg = threadIdx.x >> 5;
  const long long k = blockIdx.x * 64LL + 2 * kk;
  const long long il0 = blockIdx.y * (long long)IL;
  const bool kin = k < H1;
  float ax[IL], ay[IL];
#pragma unroll
  for (int i = 0; i < IL; ++i) { ax[i] = 0.f; ay[i] = 0.f; }
  if (kin) {
    // blockIdx.z splits the j range so that large batches (few il rows per panel) still fill the machine
    const long long jchunk = (Bk + gridDim.z - 1) / gridDim.z;
    const long long j0 = blockIdx.z * jchunk, j1 = (j0 + jchunk < Bk) ? j0 + jchunk : Bk;
    for (long long j = j0 + jg; j < j1; j += JG) {
      float sx = 0.f, sy = 0.f;
#pragma unroll
      for (int i = 0; i < IL; ++i) {
        if (il0 + i < R) {
          const T* src = dz1 + ((il0 + i) * Bk + j) * ld + k;
          float vx, vy;
          if constexpr (sizeof(T) == 2) {
            const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src));
            vx = v.x; vy = v.y;
          } else {
            const float2 v = *reinterpret_cast<const float2*>(src);
            vx = v.x; vy = v.y;
          }
          ax[i] += vx; ay[i] += vy; sx += vx; sy += vy;
        }
      }
      atomicAdd(dC + j * H1 + k, sx);
      atomicAdd(dC + j * H1 + k + 1, sy);
    }
  }
  for (int t = threadIdx.x; t < IL * 64; t += 256) sh[t >> 6][t & 63] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < IL; ++i) { atomicAdd(&sh[i][2 * kk], ax[i]); atomicAdd(&sh[i][2 * kk + 1], ay[i]); }   // 8 j groups per cell
  __syncthreads();
  for (int t = threadIdx.x; t < IL * 64; t += 256) {
    const int i = t >> 6, c = t & 63;
    if (il0 + i < R && blockIdx.x * 64LL + c < H1) atomicAdd(dA + (r0 + il0 + i) * H1 + blockIdx.x * 64LL + c, sh[i][c]);   // dA zeroed by the caller
  }
}
// out[c] = sum_r in[r, c]  (one thread per column; rows are few thousand at most)
__global__ void colsum_kernel(const float* __restrict__ in, long long R, long long C, float* __restrict__ out) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f;
  for (long long r = 0; r < R; ++r) a += in[r * C + c];
  out[c] = a;
}
// one block: out[0] = sum in[0..n)
__global__ void sum_reduce_kernel(const float* __restrict__ in, long long n, float* __restrict__ out) {
  __shared__ double sh[32];
  double a = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) a += in[i];
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i]; out[0] = static_cast<float>(t); }
}
__global__ void copy_f32_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

struct MlpDims { long long B, D, H1, H2; };
struct MlpParams { const float *W1, *b1, *W2, *b2, *W3, *b3; };
struct MlpGrads { float *dX, *dY, *dW1, *db1, *dW2, *db2, *dW3, *db3; };

long long g_mlp_max_pairs = 1LL << 20;                    // bounds the panel buffers (H, dZ2, dZ1): 5 GB fast, 10 GB strict
inline long long mlp_panel_rows(long long B) {
  const long long max_pairs = g_mlp_max_pairs;
  long long R = max_pairs / B;
  if (R < 1) R = 1;
  if (R > B) R = B;
  return R;
}

// hi*hi (+ lo*hi + hi*lo) K segments of a K-major x K-major product whose operands are [hi | lo] pairs
void kk_segments(GemmArgs& g, int sp, long long K) {
  const int kb = static_cast<int>(sp == 2 ? round_up(K, kSplitAlign) / bk() : cdiv(K, bk()));
  g.k_blocks = kb; g.seg_len = kb;
  if (sp == 2) { g.k_blocks = 3 * kb; g.a_seg[1] = kb; g.b_seg[2] = kb; }
}

int mlp_impl(const float* X, const float* Y, const MlpParams& prm, const int* sid, const MlpDims& d, int estimator, int precision,
             double* loss_out, float* S_out, const MlpGrads& gr, Bump& ws, cudaStream_t stream) {
  const long long B = d.B, D = d.D, H1 = d.H1, H2 = d.H2;
  if (B <= 0 || D <= 0 || H1 <= 0 || H2 <= 0 || (D % 8) != 0 || (H1 % 8) != 0 || (H2 % 8) != 0) return MI_ERR_BAD_ARG;
  if (H2 > mi::EpiMlpDz::kMaxTiles * mi::TILE_N) return MI_ERR_BAD_ARG;
  if (estimator != MI_EST_DV && estimator != MI_EST_INFONCE_REF && estimator != MI_EST_INFONCE_ROW) return MI_ERR_BAD_ARG;
  typedef __nv_bfloat16 bf;
  const int sp = (precision & 1) == MI_PREC_BF16_STRICT ? 2 : 1;
  const bool grads = gr.dX || gr.dY || gr.dW1 || gr.db1 || gr.dW2 || gr.db2 || gr.dW3 || gr.db3;
  const bool plan = grads || ws.dry;
  const long long Dp = round_up(D, kSplitAlign), Hp = round_up(H1, kSplitAlign), H2p = round_up(H2, kSplitAlign);
  const long long pX = sp == 2 ? 2 * Dp : D, pH = sp == 2 ? 2 * Hp : H1, pZ = sp == 2 ? 2 * H2p : H2;
  const long long eX = sp == 2 ? Dp + D : D, eH = sp == 2 ? Hp + H1 : H1, eZ = sp == 2 ? H2p + H2 : H2;   // TMA extents
  const long long R = mlp_panel_rows(B), Pmax = R * B;
  const int nt2 = static_cast<int>(cdiv(H2, mi::TILE_N));
  const long long h2_pad = static_cast<long long>(nt2) * mi::TILE_N;
  const long long rows_padded_max = cdiv(Pmax, rows_per_mblk()) * rows_per_mblk();

  bf* X16 = ws.take<bf>(B * pX); bf* Y16 = ws.take<bf>(B * pX);
  bf* W1x = ws.take<bf>(H1 * pX); bf* W1y = ws.take<bf>(H1 * pX);
  bf* W2h = ws.take<bf>(H2 * pH);
  float* A32 = ws.take<float>(B * H1); float* C32 = ws.take<float>(B * H1);
  float* b2p = ws.take<float>(h2_pad); float* w3p = ws.take<float>(h2_pad);
  float* S = S_out ? S_out : ws.take<float>(B * B);
  float* rows_r = ws.take<float>(B * 4);
  double* scal = ws.take<double>(8);
  float* lse_f = ws.take<float>(1);
  bf* Hpan = ws.take<bf>(Pmax * pH);
  float* part = ws.take<float>(mi::kColQuarters * rows_padded_max);
  float* G = plan ? ws.take<float>(B * B) : nullptr;
  bf* DZ2 = plan ? ws.take<bf>(Pmax * pZ) : nullptr;
  // dZ1 panel: bf16 in fast mode (half the bytes of the HBM-write-bound masked contraction), fp32 in strict mode
  float* DZ1 = (plan && sp == 2) ? ws.take<float>(Pmax * H1) : nullptr;
  bf* DZ1h = (plan && sp == 1) ? ws.take<bf>(Pmax * H1) : nullptr;
  float* dA32 = plan ? ws.take<float>(B * H1) : nullptr; float* dC32 = plan ? ws.take<float>(B * H1) : nullptr;
  bf* dA16 = plan ? ws.take<bf>(B * pH) : nullptr; bf* dC16 = plan ? ws.take<bf>(B * pH) : nullptr;
  float* acc_w3 = plan ? ws.take<float>(h2_pad) : nullptr; float* acc_b2 = plan ? ws.take<float>(h2_pad) : nullptr;
  float* dW2_tmp = (plan && !gr.dW2) ? ws.take<float>(H2 * H1) : nullptr;
  if (!ws.ok()) return MI_ERR_WORKSPACE;
  const size_t mk = ws.mark();
  const bool dry = ws.dry;
  if (!dry && (!X || !Y || !sid || !loss_out || !prm.W1 || !prm.b1 || !prm.W2 || !prm.b2 || !prm.W3 || !prm.b3)) return MI_ERR_BAD_ARG;

  auto split = [&](const float* in, long long ld_in, bf* out, long long pitch, long long Cp, long long Rr, long long Cc) -> int {
    split_f32_kernel<<<blocks_for(Rr * pitch, 256), 256, 0, stream>>>(in, ld_in, out, pitch, sp, Cp, Rr, Cc);
    MI_LAUNCH_CHECK("split_f32_kernel");
    return MI_OK;
  };
  if (!dry) {
    MI_TRY(split(X, D, X16, pX, Dp, B, D));
    MI_TRY(split(Y, D, Y16, pX, Dp, B, D));
    MI_TRY(split(prm.W1, 2 * D, W1x, pX, Dp, H1, D));            // W1 = [W1x | W1y]  (nn.Linear weight [H1, 2D])
    MI_TRY(split(prm.W1 + D, 2 * D, W1y, pX, Dp, H1, D));
    MI_TRY(split(prm.W2, H1, W2h, pH, Hp, H2, H1));
    pad_f32_kernel<<<blocks_for(h2_pad, 256), 256, 0, stream>>>(prm.b2, b2p, H2, h2_pad);
    MI_LAUNCH_CHECK("pad_f32_kernel");
    pad_f32_kernel<<<blocks_for(h2_pad, 256), 256, 0, stream>>>(prm.W3, w3p, H2, h2_pad);
    MI_LAUNCH_CHECK("pad_f32_kernel");
    if (sp == 2 || (H1 % bk()) != 0) MI_CUDA(cudaMemsetAsync(Hpan, 0, static_cast<size_t>(Pmax) * pH * sizeof(bf), stream));
    if (DZ2 && (sp == 2 || (H2 % bk()) != 0)) MI_CUDA(cudaMemsetAsync(DZ2, 0, static_cast<size_t>(Pmax) * pZ * sizeof(bf), stream));
  }
  // ---- layer 1: A = X W1x^T + b1, C = Y W1y^T  (fp32 out)
  for (int side = 0; side < 2; ++side) {
    GemmArgs g;
    g.a = MapSpec{side == 0 ? X16 : Y16, B, eX, pX};
    g.b = MapSpec{side == 0 ? W1x : W1y, H1, eX, pX};
    g.M = B; g.N = H1; kk_segments(g, sp, D);
    g.out_f32 = side == 0 ? A32 : C32; g.ld_out = H1;
    g.bias = side == 0 ? prm.b1 : nullptr;
    MI_TRY(run_gemm(g, ws, stream));
    ws.release(mk);
  }
  // Z2 = H W2^T over one panel, handed to an epilogue policy
  auto z2_sched = [&](long long P, Sched& sc) {
    sc.n_mblk = static_cast<int>(cdiv(P, rows_per_mblk()));
    sc.n_ntile = nt2; sc.n_split = 1; sc.n_ksplit = 1; sc.order = 0;
    single_segment(sc);
    const int kb = static_cast<int>(sp == 2 ? Hp / bk() : cdiv(H1, bk()));
    sc.k_blocks = kb; sc.seg_len = kb;
    if (sp == 2) { sc.k_blocks = 3 * kb; sc.a_seg[1] = kb; sc.b_seg[2] = kb; }
  };
  auto gen_panel = [&](long long r0, long long P) -> int {
    mlp_gen_kernel<<<blocks_for(P * (H1 >> 3), 256), 256, 0, stream>>>(A32, C32, H1, B, r0, P, Hpan, pH, Hp, sp);
    MI_LAUNCH_CHECK("mlp_gen_kernel");
    return MI_OK;
  };
  const long long n_panels = cdiv(B, R);
  if (!dry) {
    // ---- forward: logits of every pair
    for (long long r0 = 0; r0 < B; r0 += R) {
      const long long rr = (B - r0 < R) ? (B - r0) : R, P = rr * B;
      MI_TRY(gen_panel(r0, P));
      Sched sc; z2_sched(P, sc);
      mi::EpiMlpFwd::Params ep;
      ep.b2 = b2p; ep.w3 = w3p; ep.part = part; ep.rows_padded = sc.n_mblk * rows_per_mblk();
      MI_TRY(launch_engine<mi::EpiMlpFwd>(MapSpec{Hpan, P, eH, pH}, MapSpec{W2h, H2, eH, pH}, sc, ep, stream));
      mlp_logit_merge_kernel<<<blocks_for(P, 256), 256, 0, stream>>>(part, ep.rows_padded, P, prm.b3, S + r0 * B);
      MI_LAUNCH_CHECK("mlp_logit_merge_kernel");
    }
    // ---- estimator on S (same reductions and fp64 finalisation as the separable critics)
    mlp_row_stats_kernel<<<blocks_for(B * 32, 256), 256, 0, stream>>>(S, sid, B, reinterpret_cast<float4*>(rows_r));
    MI_LAUNCH_CHECK("mlp_row_stats_kernel");
    stats_reduce_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const float4*>(rows_r), static_cast<int>(B), scal, nullptr);
    MI_LAUNCH_CHECK("stats_reduce_kernel");
    loss_finalize_kernel<<<1, 32, 0, stream>>>(scal, nullptr, B, estimator, loss_out, lse_f, nullptr, nullptr);
    MI_LAUNCH_CHECK("loss_finalize_kernel");
  }
  if (!plan) return MI_OK;
  float* dW2 = gr.dW2 ? gr.dW2 : dW2_tmp;
  if (!dry) {
    const int dv_like = (estimator == MI_EST_DV || estimator == MI_EST_INFONCE_REF) ? 1 : 0;
    mlp_g_kernel<<<blocks_for(B * B, 256), 256, 0, stream>>>(S, sid, B, dv_like, lse_f, reinterpret_cast<const float4*>(rows_r), G);
    MI_LAUNCH_CHECK("mlp_g_kernel");
    MI_CUDA(cudaMemsetAsync(dC32, 0, static_cast<size_t>(B) * H1 * sizeof(float), stream));
    MI_CUDA(cudaMemsetAsync(dA32, 0, static_cast<size_t>(B) * H1 * sizeof(float), stream));
    MI_CUDA(cudaMemsetAsync(acc_w3, 0, static_cast<size_t>(h2_pad) * sizeof(float), stream));
    MI_CUDA(cudaMemsetAsync(acc_b2, 0, static_cast<size_t>(h2_pad) * sizeof(float), stream));
  }
  // ---- backward, panel by panel
  for (long long r0 = 0; r0 < B; r0 += R) {
    const long long rr = (B - r0 < R) ? (B - r0) : R, P = rr * B;
    if (!dry) {
      if (n_panels > 1) MI_TRY(gen_panel(r0, P));                 // a single panel is still resident from the forward
      Sched sc; z2_sched(P, sc);
      mi::EpiMlpDz::Params ep;
      ep.b2 = b2p; ep.w3 = w3p; ep.g = G + r0 * B; ep.rows = static_cast<int>(P); ep.cols = static_cast<int>(H2);
      ep.dz = DZ2; ep.dz_lo = sp == 2 ? DZ2 + H2p : nullptr; ep.pitch = pZ; ep.dw3 = acc_w3; ep.db2 = acc_b2;
      MI_TRY(launch_engine<mi::EpiMlpDz>(MapSpec{Hpan, P, eH, pH}, MapSpec{W2h, H2, eH, pH}, sc, ep, stream));
    }
    {   // dW2 += dZ2^T H : contraction over the panel's pairs, both operands MN-major
      GemmArgs g;
      const int kb = static_cast<int>(cdiv(P, bk()));
      g.a_mn = true; g.a = MapSpec{DZ2, P, eZ, pZ};
      g.b_mn = true; g.b = MapSpec{Hpan, P, eH, pH};
      g.M = H2; g.N = H1; g.k_blocks = kb; g.seg_len = kb;
      if (sp == 2) { g.k_blocks = 3 * kb; g.a_moff[1] = static_cast<int>(H2p); g.b_noff[2] = static_cast<int>(Hp); }
      const long long tiles = cdiv(H2, rows_per_mblk()) * cdiv(H1, mi::TILE_N);
      long long ks = cdiv(num_pairs(), tiles);
      if (ks > g.k_blocks / 4) ks = g.k_blocks / 4;
      if (ks < 1) ks = 1;
      g.ksplit = static_cast<int>(ks);
      g.out_f32 = dW2; g.ld_out = H1; g.accumulate = r0 > 0;
      MI_TRY(run_gemm(g, ws, stream));
      ws.release(mk);
    }
    {   // dZ1 = (dZ2 W2) . [H > 0] : W2 [H2, H1] read in place as the MN-major B operand
      GemmArgs g;
      g.a = MapSpec{DZ2, P, eZ, pZ};
      g.b_mn = true; g.b = MapSpec{W2h, H2, eH, pH};
      g.M = P; g.N = H1;
      const int kb = static_cast<int>(sp == 2 ? H2p / bk() : cdiv(H2, bk()));
      g.k_blocks = kb; g.seg_len = kb;
      if (sp == 2) { g.k_blocks = 3 * kb; g.a_seg[1] = kb; g.b_noff[2] = static_cast<int>(Hp); }
      g.out_f32 = DZ1; g.ld_out = H1; g.out_bf16 = DZ1h; g.ld_out16 = H1; g.relu_mask = Hpan; g.ld_mask = pH;
      if (dry) g.out_f32 = reinterpret_cast<float*>(16);
      MI_TRY(run_gemm(g, ws, stream));
      ws.release(mk);
    }
    if (!dry) {
      dim3 rg(static_cast<unsigned>(cdiv(H1, 64)), static_cast<unsigned>(cdiv(rr, 32)), 1);
      const long long blocks = static_cast<long long>(rg.x) * rg.y;
      long long jz = blocks >= num_sms() ? 1 : cdiv(4LL * num_sms(), blocks);
      if (jz > B / 64) jz = B / 64;
      rg.z = static_cast<unsigned>(jz < 1 ? 1 : jz);
      if (sp == 1) mlp_reduce_both_kernel<bf><<<rg, 256, 0, stream>>>(DZ1h, H1, H1, B, rr, r0, dA32, dC32);
      else mlp_reduce_both_kernel<float><<<rg, 256, 0, stream>>>(DZ1, H1, H1, B, rr, r0, dA32, dC32);
      MI_LAUNCH_CHECK("mlp_reduce_both_kernel");
    }
  }
  if (!dry) {
    if (gr.dW3) { copy_f32_kernel<<<blocks_for(H2, 256), 256, 0, stream>>>(acc_w3, gr.dW3, H2); MI_LAUNCH_CHECK("copy_f32_kernel"); }
    if (gr.db2) { copy_f32_kernel<<<blocks_for(H2, 256), 256, 0, stream>>>(acc_b2, gr.db2, H2); MI_LAUNCH_CHECK("copy_f32_kernel"); }
    if (gr.db3) { sum_reduce_kernel<<<1, 1024, 0, stream>>>(G, B * B, gr.db3); MI_LAUNCH_CHECK("sum_reduce_kernel"); }
    if (gr.db1) { colsum_kernel<<<blocks_for(H1, 128), 128, 0, stream>>>(dA32, B, H1, gr.db1); MI_LAUNCH_CHECK("colsum_kernel"); }
    MI_TRY(split(dA32, H1, dA16, pH, Hp, B, H1));
    MI_TRY(split(dC32, H1, dC16, pH, Hp, B, H1));
  }
  // ---- layer 1 backward
  for (int side = 0; side < 2; ++side) {
    float* dIn = side == 0 ? gr.dX : gr.dY;
    bf* dAc = side == 0 ? dA16 : dC16;
    bf* Wside = side == 0 ? W1x : W1y;
    bf* In16 = side == 0 ? X16 : Y16;
    if (dIn || dry) {   // dX = dA W1x : W1x [H1, D] read in place as the MN-major B operand
      GemmArgs g;
      g.a = MapSpec{dAc, B, eH, pH};
      g.b_mn = true; g.b = MapSpec{Wside, H1, eX, pX};
      g.M = B; g.N = D;
      const int kb = static_cast<int>(sp == 2 ? Hp / bk() : cdiv(H1, bk()));
      g.k_blocks = kb; g.seg_len = kb;
      if (sp == 2) { g.k_blocks = 3 * kb; g.a_seg[1] = kb; g.b_noff[2] = static_cast<int>(Dp); }
      g.out_f32 = dIn; g.ld_out = D;
      if (dry) { g.out_f32 = reinterpret_cast<float*>(16); }
      MI_TRY(run_gemm(g, ws, stream));
      ws.release(mk);
    }
    if (gr.dW1 || dry) {   // dW1[:, side] = dA^T X : contraction over the batch, both operands MN-major
      GemmArgs g;
      const int kb = static_cast<int>(cdiv(B, bk()));
      g.a_mn = true; g.a = MapSpec{dAc, B, eH, pH};
      g.b_mn = true; g.b = MapSpec{In16, B, eX, pX};
      g.M = H1; g.N = D; g.k_blocks = kb; g.seg_len = kb;
      if (sp == 2) { g.k_blocks = 3 * kb; g.a_moff[1] = static_cast<int>(Hp); g.b_noff[2] = static_cast<int>(Dp); }
      const long long tiles = cdiv(H1, rows_per_mblk()) * cdiv(D, mi::TILE_N);
      long long ks = cdiv(num_pairs(), tiles);
      if (ks > g.k_blocks / 4) ks = g.k_blocks / 4;
      if (ks < 1) ks = 1;
      g.ksplit = static_cast<int>(ks);
      g.out_f32 = gr.dW1 ? gr.dW1 + side * D : nullptr; g.ld_out = 2 * D;
      if (dry) { g.out_f32 = reinterpret_cast<float*>(16); }
      MI_TRY(run_gemm(g, ws, stream));
      ws.release(mk);
    }
  }
  return MI_OK;
}


