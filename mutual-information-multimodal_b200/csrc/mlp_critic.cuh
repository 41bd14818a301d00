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
//
// dv / infonce (global log-sum-exp) run as a SINGLE PASS (mlp_single_pass below): the softmax weights are formed against
// a reference logit fixed before the pass (maximum over a sample of pairs + margin), so the Z2 accumulator that gives
// the logit also gives dZ2 — Z2 is not recomputed (6 instead of 8 B^2 H1 H2 flops), [B, B] logits / gradients are not
// stored, and the dZ1 reductions happen in the epilogue of their GEMM (EpiMlpDa).  The positive pairs (weight -1/B, not
// a softmax weight) are a separate B-pair panel.  A guard that finds the reference outside the safe window repeats the
// step with the two-pass sequence above behind a device-side predicate.
// ============================================================================================
// fp32 [R, C] (pitch ld_in) -> bf16 [R, pitch]: hi in columns [0, C), lo (split == 2) in [Cp, Cp + C), zeros elsewhere
__global__ void split_f32_kernel(const float* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out, long long pitch,
                                 int split, long long Cp, long long R, long long C, const int* __restrict__ run_if) {
  MI_PRED(run_if);
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
// predicated memset (the two-pass repeat must not clear anything unless it runs)
__global__ void zero_u4_kernel(uint4* __restrict__ p, long long n16, const int* __restrict__ run_if) {
  MI_PRED(run_if);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
    p[i] = make_uint4(0u, 0u, 0u, 0u);
}
__global__ void fill_f32_kernel(float* __restrict__ p, long long n, float v) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// H[p, k] = relu(A[i, k] + C[j, k]) as bf16 hi (+ lo at column Hp), 8 columns per thread;
// pairs p = (i - r0) Bk + j of a row panel, or the positive pairs i = j = p (diag)
__global__ void mlp_gen_kernel(const float* __restrict__ A, const float* __restrict__ Cm, long long H1, long long Bk, long long r0,
                               long long P, __nv_bfloat16* __restrict__ H, long long pitch, long long Hp, int split, int diag,
                               const int* __restrict__ run_if) {
  MI_PRED(run_if);
  const long long h8 = H1 >> 3;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < P * h8; idx += (long long)gridDim.x * blockDim.x) {
    const long long p = idx / h8, k = (idx - p * h8) << 3;
    const long long i = diag ? p : r0 + p / Bk, j = diag ? p : p % Bk;
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
}
// The single pass' panel: pairs in the padded order p = il * Bp + j (j >= B: zero rows), plus the ReLU mask as bits
// (mask[p][w] bit b <=> A[i, 32 w + b] + C[j, 32 w + b] > 0, from the fp32 sum).
// A thread owns 8 columns of kGenJ text rows j (its C values stay in registers) and walks the image rows of its
// blockIdx.y slice: the A row of an image is fetched once per kGenJ pairs, C never again; 4 * wpr threads per text row
// (the 4 threads of a mask word are consecutive lanes), every 2 KB row segment of H is written by consecutive threads.
constexpr int kGenJ = 4;
__global__ void __launch_bounds__(256, 3) mlp_gen2_kernel(const float* __restrict__ A, const float* __restrict__ Cm, long long H1, long long B,
                                                       long long Bp, long long r0, long long rr, __nv_bfloat16* __restrict__ H,
                                                       long long pitch, long long Hp, int split, uint32_t* __restrict__ mask, int wpr) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long tpr = 4LL * wpr;
  const long long jg = t / tpr;
  const int k = static_cast<int>(t - jg * tpr) << 3;
  const int lane = threadIdx.x & 31;
  const bool kin = k < H1;
  float c[kGenJ][8];
  bool jin[kGenJ], real[kGenJ];
#pragma unroll
  for (int q = 0; q < kGenJ; ++q) {
    const long long j = jg * kGenJ + q;
    jin[q] = j < Bp; real[q] = kin && j < B;
#pragma unroll
    for (int e = 0; e < 8; ++e) c[q][e] = 0.f;
    if (real[q]) {
      const float4* c4 = reinterpret_cast<const float4*>(Cm + j * H1 + k);
      const float4 c0 = __ldg(c4), c1 = __ldg(c4 + 1);
      c[q][0] = c0.x; c[q][1] = c0.y; c[q][2] = c0.z; c[q][3] = c0.w; c[q][4] = c1.x; c[q][5] = c1.y; c[q][6] = c1.z; c[q][7] = c1.w;
    }
  }
  const long long per = (rr + gridDim.y - 1) / gridDim.y;
  const long long il0 = blockIdx.y * per, il1 = il0 + per < rr ? il0 + per : rr;
  for (long long il = il0; il < il1; ++il) {
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (kin) {
      const float4* a4 = reinterpret_cast<const float4*>(A + (r0 + il) * H1 + k);
      const float4 a0 = __ldg(a4), a1 = __ldg(a4 + 1);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
    }
#pragma unroll
    for (int q = 0; q < kGenJ; ++q) {
      const long long p = il * Bp + jg * kGenJ + q;
      float v[8];
      uint32_t bits = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] = real[q] ? fmaxf(a[e] + c[q][e], 0.f) : 0.f;
        bits |= (v[e] > 0.f ? 1u : 0u) << e;
      }
      uint32_t w = bits << (8 * (lane & 3));
      w |= __shfl_xor_sync(0xffffffffu, w, 1);
      w |= __shfl_xor_sync(0xffffffffu, w, 2);
      if (jin[q] && (lane & 3) == 0) mask[p * wpr + (k >> 5)] = w;
      if (jin[q] && kin) {
        uint32_t hi[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) hi[e] = ptx::pack_bf16(v[2 * e], v[2 * e + 1]);
        *reinterpret_cast<uint4*>(H + p * pitch + k) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (split == 2) {
          uint32_t lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            lo[e] = ptx::pack_bf16(v[2 * e] - __uint_as_float(hi[e] << 16), v[2 * e + 1] - __uint_as_float(hi[e] & 0xffff0000u));
          *reinterpret_cast<uint4*>(H + p * pitch + Hp + k) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
  }
}
// dst[s, :] = src[s * stride, :]  (the sampled rows of the reference-logit estimate)
__global__ void gather_rows_kernel(const float* __restrict__ src, long long stride, long long C, float* __restrict__ dst, long long n) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n * C) return;
  const long long s = idx / C, c = idx - s * C;
  dst[idx] = src[s * stride * C + c];
}
// S[r0 * Bk + p] = b3 + sum of the column-quarter partials
__global__ void mlp_logit_merge_kernel(const float* __restrict__ part, int rows_padded, long long P, const float* __restrict__ b3,
                                       float* __restrict__ S, const int* __restrict__ run_if) {
  MI_PRED(run_if);
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= P) return;
  float a = b3[0];
#pragma unroll
  for (int q = 0; q < mi::kColQuarters; ++q) a += part[(size_t)q * rows_padded + p];
  S[p] = a;
}
// one warp per row of S: {lse over the negatives, #negatives, S_ii, lse over negatives and the positive}
__global__ void mlp_row_stats_kernel(const float* __restrict__ S, const int* __restrict__ sid, long long B, float4* __restrict__ row_out,
                                     const int* __restrict__ run_if) {
  MI_PRED(run_if);
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
// the same float4 per row from the single pass' sums: lse over the negatives = ref + log(sum_j g~)
__global__ void mlp_rows_from_sums_kernel(const float* __restrict__ rowsum, const float* __restrict__ rowcnt,
                                          const float* __restrict__ diag, const float* __restrict__ ref, long long B,
                                          float4* __restrict__ row_out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= B) return;
  float s = rowsum[i];
  if (!(s >= 0.f)) s = INFINITY;                            // NaN must trip the guard, not vanish in a max
  const float cnt = rowcnt[i], d = diag[i];
  const float lse_neg = cnt > 0.f ? static_cast<float>(static_cast<double>(ref[0]) + mi::kMlpSpMargin + log(static_cast<double>(s)))
                                  : mi::neg_inf();
  const float hi = fmaxf(lse_neg, d), lo = fminf(lse_neg, d);
  row_out[i] = make_float4(lse_neg, cnt, d, hi + log1pf(expf(lo - hi)));
}
// G = dL/dS.  dv / infonce (mi_critics.py:3-23): softmax over ALL negatives, -1/B on the diagonal;
// row InfoNCE: (1/B) softmax over {positive} u negatives of the row, minus 1/B on the diagonal.
__global__ void mlp_g_kernel(const float* __restrict__ S, const int* __restrict__ sid, long long B, int dv_like,
                             const float* __restrict__ lse, const float4* __restrict__ rows, float* __restrict__ G,
                             const int* __restrict__ run_if) {
  MI_PRED(run_if);
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
                                                              long long R, long long r0, float* __restrict__ dA, float* __restrict__ dC,
                                                              const int* __restrict__ run_if) {
  MI_PRED(run_if);
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
// out[c] = sum_r in[r, c]: a block owns 32 columns, its 8 warps stride over the rows (coalesced 128 B reads)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ in, long long R, long long C, float* __restrict__ out) {
  __shared__ float sh[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long c = blockIdx.x * 32LL + lane;
  float a = 0.f;
  if (c < C) for (long long r = w; r < R; r += 8) a += in[r * C + c];
  sh[w][lane] = a;
  __syncthreads();
  if (w == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][lane];
    out[c] = t;
  }
}
// one block: out[0] = sum in[0..n)
__global__ void sum_reduce_kernel(const float* __restrict__ in, long long n, float* __restrict__ out, const int* __restrict__ run_if) {
  MI_PRED(run_if);
  __shared__ double sh[32];
  double a = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) a += in[i];
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i]; out[0] = static_cast<float>(t); }
}
// one block, after loss_finalize_kernel: total = sum of g~ over all negatives (fp64).  The softmax weights are g~ / total,
// so cfac = 1 / total takes the sums from reference units to weights without going through e^{ref - LSE}.
// Guard (loss_out[7] = 1, cfac = 0): total outside [1e-30, 1e30] (overflow of some e^{S - ref}, or every negative
// flushed to zero) while negatives exist.  db3 = sum of dL/dlogit over all pairs = total cfac - 1 (analytically 0).
__global__ void mlp_weights_kernel(const float* __restrict__ rowsum, long long n, float* __restrict__ cfac, float* __restrict__ db3,
                                   double* __restrict__ loss_out) {
  __shared__ double sh[32];
  double a = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) a += rowsum[i];
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    const bool any_neg = loss_out[3] > 0.0;
    const bool ok = t >= 1e-30 && t <= 1e30;
    const float c = (any_neg && ok) ? static_cast<float>(1.0 / t) : 0.f;
    cfac[0] = c;
    db3[0] = any_neg ? static_cast<float>(t * static_cast<double>(c) - 1.0) : -1.f;
    if (any_neg && !ok) loss_out[7] = 1.0;
  }
}
__global__ void copy_f32_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
// x *= cfac: the single pass' sums leave reference units
__global__ void mlp_rescale_kernel(float* __restrict__ x, long long n, const float* __restrict__ cfac) {
  const float c = cfac[0];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] *= c;
}
// a += z; c += z  (the positive pairs (i, i) feed row i of dA and of dC)
__global__ void mlp_add_diag_kernel(const float* __restrict__ z, long long n, float* __restrict__ a, float* __restrict__ c) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = z[i];
    a[i] += v; c[i] += v;
  }
}

struct MlpDims { long long B, D, H1, H2; };
struct MlpParams { const float *W1, *b1, *W2, *b2, *W3, *b3; };
struct MlpGrads { float *dX, *dY, *dW1, *db1, *dW2, *db2, *dW3, *db3; };

long long g_mlp_max_pairs = 1LL << 20;                    // bounds the panel buffers (H, dZ2, dZ1): 5 GB fast, 10 GB strict
std::atomic<int> g_mlp_mode{3};                           // bit 0: single pass for dv-like estimators, bit 1: fused dZ1 reductions
inline long long mlp_panel_rows(long long B, long long cols) {
  long long R = g_mlp_max_pairs / cols;
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

// memset that respects the launch predicate of this thread
int zero_async(void* p, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return MI_OK;
  if (t_run_if == nullptr) { MI_CUDA(cudaMemsetAsync(p, 0, bytes, stream)); return MI_OK; }
  const long long n16 = static_cast<long long>((bytes + 15) / 16);       // workspace blocks are 256 B aligned and padded
  zero_u4_kernel<<<blocks_capped(n16, 256), 256, 0, stream>>>(static_cast<uint4*>(p), n16, t_run_if);
  MI_LAUNCH_CHECK("zero_u4_kernel");
  return MI_OK;
}

// everything the two sequences share: dimensions, operands, scratch
struct MlpCtx {
  long long B, D, H1, H2, Dp, Hp, H2p, pX, pH, pZ, eX, eH, eZ, h2_pad;
  int sp, nt2, estimator;
  const int* sid;
  MlpParams prm;
  __nv_bfloat16 *X16, *Y16, *W1x, *W1y, *W2h, *Hpan, *DZ2, *DZ1h, *dA16, *dC16;
  float *A32, *C32, *b2p, *w3p, *S, *rows_r, *lse_f, *part, *G, *DZ1, *dA32, *dC32, *acc_w3, *acc_b2, *dW2, *db3v;
  double* scal;
  double* loss_out;
  cudaStream_t stream;
  size_t mk;
};

// Z2 = H W2^T over one panel of P pairs, handed to an epilogue policy
void mlp_z2_sched(const MlpCtx& c, long long P, Sched& sc) {
  sc.n_mblk = static_cast<int>(cdiv(P, rows_per_mblk()));
  sc.n_ntile = c.nt2; sc.n_split = 1; sc.n_ksplit = 1; sc.order = 0;
  single_segment(sc);
  const int kb = static_cast<int>(c.sp == 2 ? c.Hp / bk() : cdiv(c.H1, bk()));
  sc.k_blocks = kb; sc.seg_len = kb;
  if (c.sp == 2) { sc.k_blocks = 3 * kb; sc.a_seg[1] = kb; sc.b_seg[2] = kb; }
}
// dW2 (+)= dZ2^T H : contraction over the panel's pairs, both operands MN-major
int mlp_dw2_gemm(const MlpCtx& c, long long P, bool accumulate, Bump& ws) {
  GemmArgs g;
  const int kb = static_cast<int>(cdiv(P, bk()));
  g.a_mn = true; g.a = MapSpec{c.DZ2, P, c.eZ, c.pZ};
  g.b_mn = true; g.b = MapSpec{c.Hpan, P, c.eH, c.pH};
  g.M = c.H2; g.N = c.H1; g.k_blocks = kb; g.seg_len = kb;
  if (c.sp == 2) { g.k_blocks = 3 * kb; g.a_moff[1] = static_cast<int>(c.H2p); g.b_noff[2] = static_cast<int>(c.Hp); }
  const long long tiles = cdiv(c.H2, rows_per_mblk()) * cdiv(c.H1, mi::TILE_N);
  g.ksplit = choose_ksplit(tiles, g.k_blocks);
  g.order = 0;      // the 8 tiles of a K split stream the same rows of dZ2 / H: neighbours share them in L2
  g.out_f32 = c.dW2; g.ld_out = c.H1; g.accumulate = accumulate;
  if (ws.dry) g.out_f32 = reinterpret_cast<float*>(16);
  MI_TRY(run_gemm(g, ws, c.stream));
  ws.release(c.mk);
  return MI_OK;
}
// dZ1 = (dZ2 W2) . [H > 0] as a stored panel: W2 [H2, H1] read in place as the MN-major B operand
int mlp_dz1_gemm(const MlpCtx& c, long long P, float* out_f32, __nv_bfloat16* out_bf16, Bump& ws) {
  GemmArgs g;
  g.a = MapSpec{c.DZ2, P, c.eZ, c.pZ};
  g.b_mn = true; g.b = MapSpec{c.W2h, c.H2, c.eH, c.pH};
  g.M = P; g.N = c.H1;
  const int kb = static_cast<int>(c.sp == 2 ? c.H2p / bk() : cdiv(c.H2, bk()));
  g.k_blocks = kb; g.seg_len = kb;
  if (c.sp == 2) { g.k_blocks = 3 * kb; g.a_seg[1] = kb; g.b_noff[2] = static_cast<int>(c.Hp); }
  g.out_f32 = out_f32; g.ld_out = c.H1; g.out_bf16 = out_bf16; g.ld_out16 = c.H1; g.relu_mask = c.Hpan; g.ld_mask = c.pH;
  if (ws.dry) g.out_f32 = reinterpret_cast<float*>(16);
  MI_TRY(run_gemm(g, ws, c.stream));
  ws.release(c.mk);
  return MI_OK;
}
int mlp_reduce_panel(const MlpCtx& c, long long cols, long long rr, long long r0) {
  dim3 rg(static_cast<unsigned>(cdiv(c.H1, 64)), static_cast<unsigned>(cdiv(rr, 32)), 1);
  const long long blocks = static_cast<long long>(rg.x) * rg.y;
  long long jz = blocks >= num_sms() ? 1 : cdiv(4LL * num_sms(), blocks);
  if (jz > cols / 64) jz = cols / 64;
  rg.z = static_cast<unsigned>(jz < 1 ? 1 : jz);
  if (c.sp == 1) mlp_reduce_both_kernel<__nv_bfloat16><<<rg, 256, 0, c.stream>>>(c.DZ1h, c.H1, c.H1, cols, rr, r0, c.dA32, c.dC32, t_run_if);
  else mlp_reduce_both_kernel<float><<<rg, 256, 0, c.stream>>>(c.DZ1, c.H1, c.H1, cols, rr, r0, c.dA32, c.dC32, t_run_if);
  MI_LAUNCH_CHECK("mlp_reduce_both_kernel");
  return MI_OK;
}
int mlp_zero_accumulators(const MlpCtx& c, long long dc_rows) {
  MI_TRY(zero_async(c.dC32, static_cast<size_t>(dc_rows) * c.H1 * sizeof(float), c.stream));
  MI_TRY(zero_async(c.dA32, static_cast<size_t>(c.B) * c.H1 * sizeof(float), c.stream));
  MI_TRY(zero_async(c.acc_w3, static_cast<size_t>(c.h2_pad) * sizeof(float), c.stream));
  MI_TRY(zero_async(c.acc_b2, static_cast<size_t>(c.h2_pad) * sizeof(float), c.stream));
  return MI_OK;
}

// ---- the two-pass sequence: logits of every pair -> estimator -> recompute Z2 for the backward.  Exact for every input
//      (running-max statistics); every estimator.  All launches carry this thread's launch predicate.
int mlp_two_pass(const MlpCtx& c, bool plan, Bump& ws) {
  const long long B = c.B, H1 = c.H1, H2 = c.H2;
  const bool dry = ws.dry;
  cudaStream_t stream = c.stream;
  const long long R = mlp_panel_rows(B, B);
  const long long n_panels = cdiv(B, R);
  auto gen_panel = [&](long long r0, long long P) -> int {
    mlp_gen_kernel<<<blocks_capped(P * (H1 >> 3), 256), 256, 0, stream>>>(c.A32, c.C32, H1, B, r0, P, c.Hpan, c.pH, c.Hp, c.sp, 0, t_run_if);
    MI_LAUNCH_CHECK("mlp_gen_kernel");
    return MI_OK;
  };
  if (!dry) {
    // ---- forward: logits of every pair
    for (long long r0 = 0; r0 < B; r0 += R) {
      const long long rr = (B - r0 < R) ? (B - r0) : R, P = rr * B;
      MI_TRY(gen_panel(r0, P));
      Sched sc; mlp_z2_sched(c, P, sc);
      mi::EpiMlpFwd::Params ep;
      ep.b2 = c.b2p; ep.w3 = c.w3p; ep.part = c.part; ep.rows_padded = sc.n_mblk * rows_per_mblk();
      MI_TRY(launch_engine<mi::EpiMlpFwd>(MapSpec{c.Hpan, P, c.eH, c.pH}, MapSpec{c.W2h, H2, c.eH, c.pH}, sc, ep, stream));
      mlp_logit_merge_kernel<<<blocks_for(P, 256), 256, 0, stream>>>(c.part, ep.rows_padded, P, c.prm.b3, c.S + r0 * B, t_run_if);
      MI_LAUNCH_CHECK("mlp_logit_merge_kernel");
    }
    // ---- estimator on S (same reductions and fp64 finalisation as the separable critics)
    mlp_row_stats_kernel<<<blocks_for(B * 32, 256), 256, 0, stream>>>(c.S, c.sid, B, reinterpret_cast<float4*>(c.rows_r), t_run_if);
    MI_LAUNCH_CHECK("mlp_row_stats_kernel");
    stats_reduce_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const float4*>(c.rows_r), static_cast<int>(B), c.scal, t_run_if);
    MI_LAUNCH_CHECK("stats_reduce_kernel");
    loss_finalize_kernel<<<1, 32, 0, stream>>>(c.scal, nullptr, B, c.estimator, c.loss_out, c.lse_f, nullptr, t_run_if);
    MI_LAUNCH_CHECK("loss_finalize_kernel");
  }
  if (!plan) return MI_OK;
  if (!dry) {
    const int dv_like = (c.estimator == MI_EST_DV || c.estimator == MI_EST_INFONCE_REF) ? 1 : 0;
    mlp_g_kernel<<<blocks_for(B * B, 256), 256, 0, stream>>>(c.S, c.sid, B, dv_like, c.lse_f, reinterpret_cast<const float4*>(c.rows_r), c.G, t_run_if);
    MI_LAUNCH_CHECK("mlp_g_kernel");
    MI_TRY(mlp_zero_accumulators(c, B));
  }
  // ---- backward, panel by panel
  for (long long r0 = 0; r0 < B; r0 += R) {
    const long long rr = (B - r0 < R) ? (B - r0) : R, P = rr * B;
    if (!dry) {
      if (n_panels > 1) MI_TRY(gen_panel(r0, P));                 // a single panel is still resident from the forward
      Sched sc; mlp_z2_sched(c, P, sc);
      mi::EpiMlpDz::Params ep;
      ep.b2 = c.b2p; ep.w3 = c.w3p; ep.g = c.G + r0 * B; ep.rows = static_cast<int>(P); ep.cols = static_cast<int>(H2);
      ep.dz = c.DZ2; ep.dz_lo = c.sp == 2 ? c.DZ2 + c.H2p : nullptr; ep.pitch = c.pZ; ep.dw3 = c.acc_w3; ep.db2 = c.acc_b2;
      MI_TRY(launch_engine<mi::EpiMlpDz>(MapSpec{c.Hpan, P, c.eH, c.pH}, MapSpec{c.W2h, H2, c.eH, c.pH}, sc, ep, stream));
    }
    MI_TRY(mlp_dw2_gemm(c, P, r0 > 0, ws));
    MI_TRY(mlp_dz1_gemm(c, P, c.sp == 2 ? c.DZ1 : nullptr, c.sp == 1 ? c.DZ1h : nullptr, ws));
    if (!dry) MI_TRY(mlp_reduce_panel(c, B, rr, r0));
  }
  if (!dry) {
    sum_reduce_kernel<<<1, 1024, 0, stream>>>(c.G, B * B, c.db3v, t_run_if);
    MI_LAUNCH_CHECK("sum_reduce_kernel");
  }
  return MI_OK;
}

// reference of the single pass: the maximum logit over an n x n strided sample of the pairs (the margin is added in the epilogue)
constexpr long long kMlpRefSample = 256;
int mlp_reference_logit(const MlpCtx& c, float* As, float* Cs, float* Ssmp, float* ref) {
  const long long ns = c.B < kMlpRefSample ? c.B : kMlpRefSample, stride = c.B / ns, P = ns * ns;
  cudaStream_t stream = c.stream;
  gather_rows_kernel<<<blocks_for(ns * c.H1, 256), 256, 0, stream>>>(c.A32, stride, c.H1, As, ns);
  MI_LAUNCH_CHECK("gather_rows_kernel");
  gather_rows_kernel<<<blocks_for(ns * c.H1, 256), 256, 0, stream>>>(c.C32, stride, c.H1, Cs, ns);
  MI_LAUNCH_CHECK("gather_rows_kernel");
  mlp_gen_kernel<<<blocks_capped(P * (c.H1 >> 3), 256), 256, 0, stream>>>(As, Cs, c.H1, ns, 0, P, c.Hpan, c.pH, c.Hp, c.sp, 0, nullptr);
  MI_LAUNCH_CHECK("mlp_gen_kernel");
  Sched sc; mlp_z2_sched(c, P, sc);
  mi::EpiMlpFwd::Params ep;
  ep.b2 = c.b2p; ep.w3 = c.w3p; ep.part = c.part; ep.rows_padded = sc.n_mblk * rows_per_mblk();
  MI_TRY(launch_engine<mi::EpiMlpFwd>(MapSpec{c.Hpan, P, c.eH, c.pH}, MapSpec{c.W2h, c.H2, c.eH, c.pH}, sc, ep, stream));
  mlp_logit_merge_kernel<<<blocks_for(P, 256), 256, 0, stream>>>(c.part, ep.rows_padded, P, c.prm.b3, Ssmp, nullptr);
  MI_LAUNCH_CHECK("mlp_logit_merge_kernel");
  max_reduce_kernel<<<1, 1024, 0, stream>>>(Ssmp, P, ref, 0.f, nullptr);
  MI_LAUNCH_CHECK("max_reduce_kernel");
  return MI_OK;
}

// image rows walked per unit of the fused dZ1 kernel: the unit count should fill whole rounds of the CTA pairs
void mlp_da_chunks(long long rr, long long per_chunk_units, int& chunk_len, int& n_chunks) {
  const long long U = num_pairs();
  const long long min_len = rr < 32 ? rr : 32;          // the flush of a unit's running sums is amortised over >= 32 tiles
  double best = -1.0;
  chunk_len = static_cast<int>(rr); n_chunks = 1;
  const long long nc_max = rr / min_len < 1 ? 1 : rr / min_len;         // full-length chunks stay >= min_len
  for (long long nc = 1; nc <= nc_max; ++nc) {
    const long long len = cdiv(rr, nc), real_nc = cdiv(rr, len), units = per_chunk_units * real_nc;
    const double eff = static_cast<double>(units) / static_cast<double>(cdiv(units, U) * U);
    if (eff > best + 1e-9) { best = eff; chunk_len = static_cast<int>(len); n_chunks = static_cast<int>(real_nc); }
  }
}

// ---- the single pass (dv / infonce): see the header of this file
struct MlpSpBuf { float *As, *Cs, *Ssmp, *ref, *cfac, *rowsum, *rowcnt, *diag, *gd, *DZ1d; uint32_t* mask; int* flag; double* guard_a; };
int mlp_single_pass(const MlpCtx& c, const MlpSpBuf& sb, float* S_out, Bump& ws) {
  const long long B = c.B, H1 = c.H1, H2 = c.H2;
  const bool dry = ws.dry;
  cudaStream_t stream = c.stream;
  const long long Bp = round_up(B, rows_per_mblk());
  const long long R = mlp_panel_rows(B, Bp);
  const int wpr = static_cast<int>(round_up(cdiv(H1, 32), 2));          // even: the epilogue reads word pairs
  const bool fused_da = (g_mlp_mode.load(std::memory_order_relaxed) & 2) != 0;
  if (!dry) {
    MI_TRY(mlp_reference_logit(c, sb.As, sb.Cs, sb.Ssmp, sb.ref));
    MI_TRY(mlp_zero_accumulators(c, Bp));
    MI_CUDA(cudaMemsetAsync(sb.rowsum, 0, static_cast<size_t>(2 * B) * sizeof(float), stream));      // rowsum | rowcnt
  }
  for (long long r0 = 0; r0 < B; r0 += R) {
    const long long rr = (B - r0 < R) ? (B - r0) : R, P = rr * Bp;
    if (!dry) {
      {
        dim3 gg(blocks_for(cdiv(Bp, kGenJ) * 4 * wpr, 256), 1, 1);
        long long ny = cdiv(16LL * num_sms(), gg.x);              // enough blocks in flight to saturate the HBM writes
        if (ny > rr) ny = rr;
        gg.y = static_cast<unsigned>(ny < 1 ? 1 : ny);
        mlp_gen2_kernel<<<gg, 256, 0, stream>>>(c.A32, c.C32, H1, B, Bp, r0, rr, c.Hpan, c.pH, c.Hp, c.sp, sb.mask, wpr);
        MI_LAUNCH_CHECK("mlp_gen2_kernel");
      }
      Sched sc; mlp_z2_sched(c, P, sc);
      mi::EpiMlpSp::Params ep;
      ep.b2 = c.b2p; ep.w3 = c.w3p; ep.b3 = c.prm.b3; ep.ref = sb.ref; ep.sid = c.sid;
      ep.B = static_cast<int>(B); ep.Bp = static_cast<int>(Bp); ep.r0 = static_cast<int>(r0); ep.rr = static_cast<int>(rr);
      ep.cols = static_cast<int>(H2);
      ep.dz = c.DZ2; ep.dz_lo = c.sp == 2 ? c.DZ2 + c.H2p : nullptr; ep.pitch = c.pZ; ep.dw3 = c.acc_w3; ep.db2 = c.acc_b2;
      ep.rowsum = sb.rowsum; ep.rowcnt = sb.rowcnt; ep.diag_out = sb.diag; ep.S_out = S_out;
      MI_TRY(launch_engine<mi::EpiMlpSp>(MapSpec{c.Hpan, P, c.eH, c.pH}, MapSpec{c.W2h, H2, c.eH, c.pH}, sc, ep, stream));
    }
    MI_TRY(mlp_dw2_gemm(c, P, r0 > 0, ws));
    if (fused_da) {
      if (!dry) {
        Sched sc;
        single_segment(sc);
        sc.order = 2;
        sc.n_ntile = static_cast<int>(cdiv(H1, mi::TILE_N));
        sc.grp = static_cast<int>(Bp / rows_per_mblk()); sc.n_il = static_cast<int>(rr);
        mlp_da_chunks(rr, static_cast<long long>(sc.n_ntile) * sc.grp, sc.chunk_len, sc.n_chunks);
        sc.n_mblk = sc.grp * sc.n_il; sc.n_split = 1; sc.n_ksplit = 1;
        const int kb = static_cast<int>(c.sp == 2 ? c.H2p / bk() : cdiv(H2, bk()));
        sc.k_blocks = kb; sc.seg_len = kb;
        if (c.sp == 2) { sc.k_blocks = 3 * kb; sc.a_seg[1] = kb; sc.b_noff[2] = static_cast<int>(c.Hp); }
        mi::EpiMlpDa::Params ep;
        ep.mask = sb.mask; ep.wpr = wpr; ep.B = static_cast<int>(B); ep.Bp = static_cast<int>(Bp);
        ep.r0 = static_cast<int>(r0); ep.rr = static_cast<int>(rr); ep.cols = static_cast<int>(H1);
        ep.dA = c.dA32; ep.dC = c.dC32;
        MI_TRY((launch_engine<mi::EpiMlpDa, false, true>(MapSpec{c.DZ2, P, c.eZ, c.pZ}, MapSpec{c.W2h, H2, c.eH, c.pH}, sc, ep, stream)));
      }
    } else {
      MI_TRY(mlp_dz1_gemm(c, P, c.sp == 2 ? c.DZ1 : nullptr, c.sp == 1 ? c.DZ1h : nullptr, ws));
      if (!dry) MI_TRY(mlp_reduce_panel(c, Bp, rr, r0));
    }
  }
  if (!dry) {
    // ---- estimator: row statistics from the sums, the shared fp64 finalisation (its guard checks |ref - LSE|)
    mlp_rows_from_sums_kernel<<<blocks_for(B, 256), 256, 0, stream>>>(sb.rowsum, sb.rowcnt, sb.diag, sb.ref, B, reinterpret_cast<float4*>(c.rows_r));
    MI_LAUNCH_CHECK("mlp_rows_from_sums_kernel");
    stats_reduce_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const float4*>(c.rows_r), static_cast<int>(B), c.scal, nullptr);
    MI_LAUNCH_CHECK("stats_reduce_kernel");
    loss_finalize_kernel<<<1, 32, 0, stream>>>(c.scal, nullptr, B, c.estimator, c.loss_out, c.lse_f, nullptr, nullptr);
    MI_LAUNCH_CHECK("loss_finalize_kernel");
    mlp_weights_kernel<<<1, 1024, 0, stream>>>(sb.rowsum, B, sb.cfac, c.db3v, c.loss_out);
    MI_LAUNCH_CHECK("mlp_weights_kernel");
    // ---- reference units -> softmax weights
    float* bufs[5] = {c.dW2, c.dA32, c.dC32, c.acc_w3, c.acc_b2};
    const long long lens[5] = {H2 * H1, B * H1, B * H1, c.h2_pad, c.h2_pad};
    for (int t = 0; t < 5; ++t) {
      mlp_rescale_kernel<<<blocks_capped(lens[t], 256), 256, 0, stream>>>(bufs[t], lens[t], sb.cfac);
      MI_LAUNCH_CHECK("mlp_rescale_kernel");
    }
    // ---- the positive pairs (i, i): dL/dlogit = -1/B (mi_critics.py:6,18), one panel of B pairs
    mlp_gen_kernel<<<blocks_capped(B * (H1 >> 3), 256), 256, 0, stream>>>(c.A32, c.C32, H1, B, 0, B, c.Hpan, c.pH, c.Hp, c.sp, 1, nullptr);
    MI_LAUNCH_CHECK("mlp_gen_kernel");
    fill_f32_kernel<<<blocks_for(B, 256), 256, 0, stream>>>(sb.gd, B, -1.f / static_cast<float>(B));
    MI_LAUNCH_CHECK("fill_f32_kernel");
    Sched sc; mlp_z2_sched(c, B, sc);
    mi::EpiMlpDz::Params ep;
    ep.b2 = c.b2p; ep.w3 = c.w3p; ep.g = sb.gd; ep.rows = static_cast<int>(B); ep.cols = static_cast<int>(H2);
    ep.dz = c.DZ2; ep.dz_lo = c.sp == 2 ? c.DZ2 + c.H2p : nullptr; ep.pitch = c.pZ; ep.dw3 = c.acc_w3; ep.db2 = c.acc_b2;
    MI_TRY(launch_engine<mi::EpiMlpDz>(MapSpec{c.Hpan, B, c.eH, c.pH}, MapSpec{c.W2h, H2, c.eH, c.pH}, sc, ep, stream));
  }
  MI_TRY(mlp_dw2_gemm(c, B, true, ws));
  MI_TRY(mlp_dz1_gemm(c, B, sb.DZ1d, nullptr, ws));
  if (!dry) {
    mlp_add_diag_kernel<<<blocks_capped(B * H1, 256), 256, 0, stream>>>(sb.DZ1d, B * H1, c.dA32, c.dC32);
    MI_LAUNCH_CHECK("mlp_add_diag_kernel");
  }
  return MI_OK;
}

int mlp_impl(const float* X, const float* Y, const MlpParams& prm, const int* sid, const MlpDims& d, int estimator, int precision,
             double* loss_out, float* S_out, const MlpGrads& gr, Bump& ws, cudaStream_t stream) {
  const long long B = d.B, D = d.D, H1 = d.H1, H2 = d.H2;
  if (B <= 0 || D <= 0 || H1 <= 0 || H2 <= 0 || (D % 8) != 0 || (H1 % 8) != 0 || (H2 % 8) != 0) return MI_ERR_BAD_ARG;
  if (H2 > mi::EpiMlpDz::kMaxTiles * mi::TILE_N) return MI_ERR_BAD_ARG;
  if (estimator != MI_EST_DV && estimator != MI_EST_INFONCE_REF && estimator != MI_EST_INFONCE_ROW) return MI_ERR_BAD_ARG;
  typedef __nv_bfloat16 bf;
  const bool grads = gr.dX || gr.dY || gr.dW1 || gr.db1 || gr.dW2 || gr.db2 || gr.dW3 || gr.db3;
  const bool plan = grads || ws.dry;
  const bool dv_like = estimator == MI_EST_DV || estimator == MI_EST_INFONCE_REF;
  // the workspace is planned for both sequences (the size query does not know the estimator's path)
  const bool single = dv_like && plan && (g_mlp_mode.load(std::memory_order_relaxed) & 1) != 0;
  MlpCtx c;
  c.B = B; c.D = D; c.H1 = H1; c.H2 = H2; c.estimator = estimator; c.sid = sid; c.prm = prm; c.stream = stream; c.loss_out = loss_out;
  c.sp = (precision & 1) == MI_PREC_BF16_STRICT ? 2 : 1;
  const int sp = c.sp;
  c.Dp = round_up(D, kSplitAlign); c.Hp = round_up(H1, kSplitAlign); c.H2p = round_up(H2, kSplitAlign);
  c.pX = sp == 2 ? 2 * c.Dp : D; c.pH = sp == 2 ? 2 * c.Hp : H1; c.pZ = sp == 2 ? 2 * c.H2p : H2;
  c.eX = sp == 2 ? c.Dp + D : D; c.eH = sp == 2 ? c.Hp + H1 : H1; c.eZ = sp == 2 ? c.H2p + H2 : H2;   // TMA extents
  const long long Bp = round_up(B, rows_per_mblk());
  const long long Pmax1 = mlp_panel_rows(B, B) * B, Pmax2 = mlp_panel_rows(B, Bp) * Bp;
  const long long ns_ref = B < kMlpRefSample ? B : kMlpRefSample;        // the reference-logit sample is a panel too
  long long Pmax = Pmax1 > Pmax2 ? Pmax1 : Pmax2;
  if (Pmax < ns_ref * ns_ref) Pmax = ns_ref * ns_ref;
  c.nt2 = static_cast<int>(cdiv(H2, mi::TILE_N));
  c.h2_pad = static_cast<long long>(c.nt2) * mi::TILE_N;
  const long long rows_padded_max = cdiv(Pmax, rows_per_mblk()) * rows_per_mblk();
  const long long pX = c.pX, pH = c.pH, pZ = c.pZ;

  c.X16 = ws.take<bf>(B * pX); c.Y16 = ws.take<bf>(B * pX);
  c.W1x = ws.take<bf>(H1 * pX); c.W1y = ws.take<bf>(H1 * pX);
  c.W2h = ws.take<bf>(H2 * pH);
  c.A32 = ws.take<float>(B * H1); c.C32 = ws.take<float>(B * H1);
  c.b2p = ws.take<float>(c.h2_pad); c.w3p = ws.take<float>(c.h2_pad);
  c.S = S_out ? S_out : ws.take<float>(B * B);
  c.rows_r = ws.take<float>(B * 4);
  c.scal = ws.take<double>(8);
  c.lse_f = ws.take<float>(1);
  c.db3v = ws.take<float>(1);
  c.Hpan = ws.take<bf>(Pmax * pH);
  c.part = ws.take<float>(mi::kColQuarters * rows_padded_max);
  c.G = plan ? ws.take<float>(B * B) : nullptr;
  c.DZ2 = plan ? ws.take<bf>(Pmax * pZ) : nullptr;
  // dZ1 panel: bf16 in fast mode (half the bytes of the HBM-write-bound masked contraction), fp32 in strict mode
  c.DZ1 = (plan && sp == 2) ? ws.take<float>(Pmax * H1) : nullptr;
  c.DZ1h = (plan && sp == 1) ? ws.take<bf>(Pmax * H1) : nullptr;
  c.dA32 = plan ? ws.take<float>(B * H1) : nullptr; c.dC32 = plan ? ws.take<float>(Bp * H1) : nullptr;
  c.dA16 = plan ? ws.take<bf>(B * pH) : nullptr; c.dC16 = plan ? ws.take<bf>(B * pH) : nullptr;
  c.acc_w3 = plan ? ws.take<float>(c.h2_pad) : nullptr; c.acc_b2 = plan ? ws.take<float>(c.h2_pad) : nullptr;
  float* dW2_tmp = (plan && !gr.dW2) ? ws.take<float>(H2 * H1) : nullptr;
  MlpSpBuf sb{};
  if (plan) {
    const long long ns = ns_ref;
    sb.As = ws.take<float>(ns * H1); sb.Cs = ws.take<float>(ns * H1); sb.Ssmp = ws.take<float>(ns * ns);
    sb.ref = ws.take<float>(1); sb.cfac = ws.take<float>(1);
    sb.rowsum = ws.take<float>(2 * B); sb.rowcnt = sb.rowsum ? sb.rowsum + B : nullptr;
    sb.diag = ws.take<float>(B); sb.gd = ws.take<float>(B);
    sb.DZ1d = ws.take<float>(B * H1);
    sb.mask = ws.take<uint32_t>(Pmax2 * round_up(cdiv(H1, 32), 2));
    sb.flag = ws.take<int>(1); sb.guard_a = ws.take<double>(1);
  }
  if (!ws.ok()) return MI_ERR_WORKSPACE;
  c.mk = ws.mark();
  c.dW2 = gr.dW2 ? gr.dW2 : dW2_tmp;
  const size_t mk = c.mk;
  const bool dry = ws.dry;
  if (!dry && (!X || !Y || !sid || !loss_out || !prm.W1 || !prm.b1 || !prm.W2 || !prm.b2 || !prm.W3 || !prm.b3)) return MI_ERR_BAD_ARG;

  auto split = [&](const float* in, long long ld_in, bf* out, long long pitch, long long Cp, long long Rr, long long Cc) -> int {
    split_f32_kernel<<<blocks_for(Rr * pitch, 256), 256, 0, stream>>>(in, ld_in, out, pitch, sp, Cp, Rr, Cc, nullptr);
    MI_LAUNCH_CHECK("split_f32_kernel");
    return MI_OK;
  };
  if (!dry) {
    MI_TRY(split(X, D, c.X16, pX, c.Dp, B, D));
    MI_TRY(split(Y, D, c.Y16, pX, c.Dp, B, D));
    MI_TRY(split(prm.W1, 2 * D, c.W1x, pX, c.Dp, H1, D));            // W1 = [W1x | W1y]  (nn.Linear weight [H1, 2D])
    MI_TRY(split(prm.W1 + D, 2 * D, c.W1y, pX, c.Dp, H1, D));
    MI_TRY(split(prm.W2, H1, c.W2h, pH, c.Hp, H2, H1));
    pad_f32_kernel<<<blocks_for(c.h2_pad, 256), 256, 0, stream>>>(prm.b2, c.b2p, H2, c.h2_pad);
    MI_LAUNCH_CHECK("pad_f32_kernel");
    pad_f32_kernel<<<blocks_for(c.h2_pad, 256), 256, 0, stream>>>(prm.W3, c.w3p, H2, c.h2_pad);
    MI_LAUNCH_CHECK("pad_f32_kernel");
    // the gaps between the hi and lo halves / beyond a ragged K are never written by the generators
    if (sp == 2 || (H1 % bk()) != 0) MI_CUDA(cudaMemsetAsync(c.Hpan, 0, static_cast<size_t>(Pmax) * pH * sizeof(bf), stream));
    if (c.DZ2 && (sp == 2 || (H2 % bk()) != 0)) MI_CUDA(cudaMemsetAsync(c.DZ2, 0, static_cast<size_t>(Pmax) * pZ * sizeof(bf), stream));
  }
  // ---- layer 1: A = X W1x^T + b1, C = Y W1y^T  (fp32 out)
  for (int side = 0; side < 2; ++side) {
    GemmArgs g;
    g.a = MapSpec{side == 0 ? c.X16 : c.Y16, B, c.eX, pX};
    g.b = MapSpec{side == 0 ? c.W1x : c.W1y, H1, c.eX, pX};
    g.M = B; g.N = H1; kk_segments(g, sp, D);
    g.out_f32 = side == 0 ? c.A32 : c.C32; g.ld_out = H1;
    g.bias = side == 0 ? prm.b1 : nullptr;
    MI_TRY(run_gemm(g, ws, stream));
    ws.release(mk);
  }
  if (single || dry) {
    MI_TRY(mlp_single_pass(c, sb, S_out, ws));
    if (!dry) {
      guard_to_flag_kernel<<<1, 1, 0, stream>>>(loss_out, sb.flag, sb.guard_a);
      MI_LAUNCH_CHECK("guard_to_flag_kernel");
    }
    {   // the two-pass sequence repeats the step if (and only if) the guard tripped
      PredGuard pg(dry ? nullptr : sb.flag);
      MI_TRY(mlp_two_pass(c, plan, ws));
    }
    if (!dry) {
      guard_report_kernel<<<1, 1, 0, stream>>>(sb.flag, sb.guard_a, loss_out);
      MI_LAUNCH_CHECK("guard_report_kernel");
    }
  } else {
    MI_TRY(mlp_two_pass(c, plan, ws));
  }
  if (!plan) return MI_OK;
  if (!dry) {
    if (gr.dW3) { copy_f32_kernel<<<blocks_for(H2, 256), 256, 0, stream>>>(c.acc_w3, gr.dW3, H2); MI_LAUNCH_CHECK("copy_f32_kernel"); }
    if (gr.db2) { copy_f32_kernel<<<blocks_for(H2, 256), 256, 0, stream>>>(c.acc_b2, gr.db2, H2); MI_LAUNCH_CHECK("copy_f32_kernel"); }
    if (gr.db3) { copy_f32_kernel<<<1, 32, 0, stream>>>(c.db3v, gr.db3, 1); MI_LAUNCH_CHECK("copy_f32_kernel"); }
    if (gr.db1) { colsum_kernel<<<blocks_for(H1, 32), 256, 0, stream>>>(c.dA32, B, H1, gr.db1); MI_LAUNCH_CHECK("colsum_kernel"); }
    MI_TRY(split(c.dA32, H1, c.dA16, pH, c.Hp, B, H1));
    MI_TRY(split(c.dC32, H1, c.dC16, pH, c.Hp, B, H1));
  }
  // ---- layer 1 backward
  for (int side = 0; side < 2; ++side) {
    float* dIn = side == 0 ? gr.dX : gr.dY;
    bf* dAc = side == 0 ? c.dA16 : c.dC16;
    bf* Wside = side == 0 ? c.W1x : c.W1y;
    bf* In16 = side == 0 ? c.X16 : c.Y16;
    if (dIn || dry) {   // dX = dA W1x : W1x [H1, D] read in place as the MN-major B operand
      GemmArgs g;
      g.a = MapSpec{dAc, B, c.eH, pH};
      g.b_mn = true; g.b = MapSpec{Wside, H1, c.eX, pX};
      g.M = B; g.N = D;
      const int kb = static_cast<int>(sp == 2 ? c.Hp / bk() : cdiv(H1, bk()));
      g.k_blocks = kb; g.seg_len = kb;
      if (sp == 2) { g.k_blocks = 3 * kb; g.a_seg[1] = kb; g.b_noff[2] = static_cast<int>(c.Dp); }
      g.out_f32 = dIn; g.ld_out = D;
      if (dry) { g.out_f32 = reinterpret_cast<float*>(16); }
      MI_TRY(run_gemm(g, ws, stream));
      ws.release(mk);
    }
    if (gr.dW1 || dry) {   // dW1[:, side] = dA^T X : contraction over the batch, both operands MN-major
      GemmArgs g;
      const int kb = static_cast<int>(cdiv(B, bk()));
      g.a_mn = true; g.a = MapSpec{dAc, B, c.eH, pH};
      g.b_mn = true; g.b = MapSpec{In16, B, c.eX, pX};
      g.M = H1; g.N = D; g.k_blocks = kb; g.seg_len = kb;
      if (sp == 2) { g.k_blocks = 3 * kb; g.a_moff[1] = static_cast<int>(c.Hp); g.b_noff[2] = static_cast<int>(c.Dp); }
      const long long tiles = cdiv(H1, rows_per_mblk()) * cdiv(D, mi::TILE_N);
      g.ksplit = choose_ksplit(tiles, g.k_blocks);
      g.out_f32 = gr.dW1 ? gr.dW1 + side * D : nullptr; g.ld_out = 2 * D;
      if (dry) { g.out_f32 = reinterpret_cast<float*>(16); }
      MI_TRY(run_gemm(g, ws, stream));
      ws.release(mk);
    }
  }
  return MI_OK;
}
