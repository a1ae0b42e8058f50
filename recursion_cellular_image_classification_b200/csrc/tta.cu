// tta.cu — kernel family 4: softmax + view averaging, plate-group mask, rescale, greedy assignment.
//
// Restates reference cell_classifier/test.py:27 (softmax), :42-45 (plate-group mask), :34-39 (rescale)
// and :48-56 (the greedy one-class-per-well loop).  The assignment must be BIT-EXACT, and the loop
// divides every row by its float32 sum between picks, so the float32 row sum has to be numpy's:
// np.sum(axis=1) on a C-contiguous float32 matrix is pairwise summation with 8 interleaved
// accumulators in blocks of <=128 (numpy/core/src/umath/loops_utils.h.src, *_pairwise_sum).  The host
// turns the recursion for a given row length into a flat plan (leaves + an RPN combine program) and
// the device replays it with the same association order; IEEE add and divide do the rest.
//
// The greedy loop is a persistent cooperative kernel: the N x C matrix lives in shared memory, a few
// rows per CTA; each of the N picks costs one grid-wide barrier.
#include <cooperative_groups.h>
#include <math.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace rxb {

constexpr int kMaxLeaves = 160;   // row length up to ~10k
constexpr int kMaxProg = 2 * kMaxLeaves;

struct PairwisePlan {
  int n_leaves;
  int n_prog;
  int C;
  uint16_t leaf_off[kMaxLeaves];
  uint8_t leaf_len[kMaxLeaves];   // <= 128
  uint8_t prog[kMaxProg];         // 0 = push next leaf, 1 = add top two (left + right)
};

static void plan_rec(PairwisePlan& p, int off, int n) {
  if (n <= 128) {
    p.leaf_off[p.n_leaves] = (uint16_t)off;
    p.leaf_len[p.n_leaves] = (uint8_t)n;
    ++p.n_leaves;
    p.prog[p.n_prog++] = 0;
    return;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  plan_rec(p, off, n2);
  plan_rec(p, off + n2, n - n2);
  p.prog[p.n_prog++] = 1;
}

static int make_plan(PairwisePlan& p, int C) {
  p.n_leaves = 0;
  p.n_prog = 0;
  p.C = C;
  // leaves are >= 64 once the row is split, so this bounds the plan
  if (C <= 0 || C > 64 * kMaxLeaves / 2 || C > 65535) return -1;
  plan_rec(p, 0, C);
  return 0;
}

// Sum of row[0..C) with numpy's association order.  Executed by one full warp; result in all lanes.
// `leaf_sums` is a per-warp shared scratch of kMaxLeaves floats.
__device__ __forceinline__ float pairwise_row_sum(const float* row, const PairwisePlan& p, float* leaf_sums,
                                                  int lane) {
  const int grp = lane >> 3, j = lane & 7;
  for (int l0 = 0; l0 < p.n_leaves; l0 += 4) {
    const int l = l0 + grp;
    float res = 0.f;
    const bool active = l < p.n_leaves;
    int off = 0, len = 0;
    if (active) {
      off = p.leaf_off[l];
      len = p.leaf_len[l];
    }
    if (active && len >= 8) {
      float r = row[off + j];
      const int body = len - (len & 7);
      for (int i = 8; i < body; i += 8) r = __fadd_rn(r, row[off + i + j]);
      res = r;
    }
    // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) : all lanes take part in the shuffles
    float t = __fadd_rn(res, __shfl_down_sync(0xffffffffu, res, 1));
    float u = __fadd_rn(t, __shfl_down_sync(0xffffffffu, t, 2));
    float w = __fadd_rn(u, __shfl_down_sync(0xffffffffu, u, 4));
    if (active && j == 0) {
      if (len >= 8) {
        for (int i = len - (len & 7); i < len; ++i) w = __fadd_rn(w, row[off + i]);
      } else {
        w = -0.0f;
        for (int i = 0; i < len; ++i) w = __fadd_rn(w, row[off + i]);
      }
      leaf_sums[l] = w;
    }
  }
  __syncwarp();
  float total = 0.f;
  if (lane == 0) {
    float stack[24];
    int sp = 0, next = 0;
    for (int i = 0; i < p.n_prog; ++i) {
      if (p.prog[i] == 0) {
        stack[sp++] = leaf_sums[next++];
      } else {
        float b = stack[--sp];
        float a = stack[--sp];
        stack[sp++] = __fadd_rn(a, b);
      }
    }
    total = stack[0];
  }
  total = __shfl_sync(0xffffffffu, total, 0);
  __syncwarp();
  return total;
}

// rescale one row in place (test.py:34-39): row /= sum, unless sum == 0.
__device__ __forceinline__ void rescale_row(float* row, const PairwisePlan& p, float* leaf_sums, int lane) {
  float s = pairwise_row_sum(row, p, leaf_sums, lane);
  if (s == 0.f) s = 1.f;
  for (int c = lane; c < p.C; c += 32) row[c] = __fdiv_rn(row[c], s);
  __syncwarp();
}

// first-index argmax of a row (np.argmax), by one warp.
__device__ __forceinline__ void row_argmax(const float* row, int C, int lane, float& best_v, int& best_c) {
  float v = -INFINITY;
  int c = 0x7fffffff;
  for (int i = lane; i < C; i += 32) {
    float x = row[i];
    if (x > v || c == 0x7fffffff) {  // strict > keeps the first index within a lane (indices ascend)
      v = x;
      c = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oc = __shfl_xor_sync(0xffffffffu, c, o);
    if (ov > v || (ov == v && oc < c)) {
      v = ov;
      c = oc;
    }
  }
  best_v = v;
  best_c = c;
}

// ------------------------------------------------------------------------------------------------
// softmax over classes per (view, row), mean over views, plate-group mask, rescale.  One CTA per row.
constexpr int kTtaThreads = 256;

__global__ void __launch_bounds__(kTtaThreads)
tta_softmax_avg_mask_kernel(const float* __restrict__ logits, int V, int N, const int32_t* __restrict__ plate,
                            const int32_t* __restrict__ group_col, float* __restrict__ probs,
                            const PairwisePlan plan, int apply_softmax) {
  extern __shared__ float smem[];
  float* row = smem;                      // C floats
  float* leaf_sums = smem + plan.C;       // kMaxLeaves
  __shared__ float red[kTtaThreads / 32];
  __shared__ float bc;
  const int n = blockIdx.x;
  const int C = plan.C;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  for (int c = threadIdx.x; c < C; c += kTtaThreads) row[c] = 0.f;
  __syncthreads();
  if (apply_softmax) {
    for (int v = 0; v < V; ++v) {
      const float* x = logits + ((long long)v * N + n) * C;
      float m = -INFINITY;
      for (int c = threadIdx.x; c < C; c += kTtaThreads) m = fmaxf(m, x[c]);
      m = warp_max(m);
      if (lane == 0) red[wid] = m;
      __syncthreads();
      if (threadIdx.x == 0) {
        float mm = red[0];
        for (int i = 1; i < kTtaThreads / 32; ++i) mm = fmaxf(mm, red[i]);
        bc = mm;
      }
      __syncthreads();
      m = bc;
      float s = 0.f;
      for (int c = threadIdx.x; c < C; c += kTtaThreads) s += expf(x[c] - m);
      s = warp_sum(s);
      __syncthreads();
      if (lane == 0) red[wid] = s;
      __syncthreads();
      if (threadIdx.x == 0) {
        float ss = 0.f;
        for (int i = 0; i < kTtaThreads / 32; ++i) ss += red[i];
        bc = ss;
      }
      __syncthreads();
      s = bc;
      for (int c = threadIdx.x; c < C; c += kTtaThreads) row[c] += __fdiv_rn(expf(x[c] - m), s);
      __syncthreads();
    }
    if (V > 1) {
      const float fv = (float)V;
      for (int c = threadIdx.x; c < C; c += kTtaThreads) row[c] = __fdiv_rn(row[c], fv);
    }
  } else {
    const float* x = probs + (long long)n * C;
    for (int c = threadIdx.x; c < C; c += kTtaThreads) row[c] = x[c];
  }
  __syncthreads();
  if (plate != nullptr) {
    const int pl = plate[n];
    for (int c = threadIdx.x; c < C; c += kTtaThreads)
      if (group_col[c] != pl) row[c] = 0.f;
  }
  __syncthreads();
  if (wid == 0) rescale_row(row, plan, leaf_sums, lane);
  __syncthreads();
  float* out = probs + (long long)n * C;
  for (int c = threadIdx.x; c < C; c += kTtaThreads) out[c] = row[c];
}

// ------------------------------------------------------------------------------------------------
// The greedy loop of test.py:48-56.
struct __align__(16) Cand {
  float v;
  int r;
  int c;
  int pad;
};
constexpr int kGreedyThreads = 256;  // 8 warps

__global__ void __launch_bounds__(kGreedyThreads)
greedy_assign_kernel(const float* __restrict__ preds_in, int N, int rows_per_cta, int32_t* __restrict__ result,
                     Cand* __restrict__ cand /*[2][gridDim.x]*/, const PairwisePlan plan) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ float smem[];
  const int C = plan.C;
  float* rows = smem;                                        // rows_per_cta * C
  float* leaf_all = smem + (size_t)rows_per_cta * C;         // 8 warps * kMaxLeaves
  __shared__ float s_best_v[kGreedyThreads / 32];
  __shared__ int s_best_r[kGreedyThreads / 32], s_best_c[kGreedyThreads / 32];
  __shared__ int s_pick_r, s_pick_c;

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nwarps = kGreedyThreads / 32;
  float* leaf_sums = leaf_all + wid * kMaxLeaves;
  const int row0 = blockIdx.x * rows_per_cta;
  const int my_rows = max(0, min(rows_per_cta, N - row0));

  for (int i = threadIdx.x; i < my_rows * C; i += kGreedyThreads) rows[i] = preds_in[(long long)row0 * C + i];
  for (int r = threadIdx.x; r < my_rows; r += kGreedyThreads) result[row0 + r] = 0;  // np.zeros
  __syncthreads();

  for (int it = 0; it < N; ++it) {
    // 1. best (value, row, col) among this CTA's rows: max value, then smallest row
    float bv = -INFINITY;
    int br = 0x7fffffff, bc = 0;
    for (int r = wid; r < my_rows; r += nwarps) {
      float v;
      int c;
      row_argmax(rows + (size_t)r * C, C, lane, v, c);
      if (v > bv || br == 0x7fffffff) {  // rows ascend within a warp: strict > keeps the first
        bv = v;
        br = row0 + r;
        bc = c;
      }
    }
    if (lane == 0) {
      s_best_v[wid] = bv;
      s_best_r[wid] = br;
      s_best_c[wid] = bc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = s_best_v[0];
      int r = s_best_r[0], c = s_best_c[0];
      for (int w = 1; w < nwarps; ++w) {
        if (s_best_r[w] == 0x7fffffff) continue;
        if (r == 0x7fffffff || s_best_v[w] > v || (s_best_v[w] == v && s_best_r[w] < r)) {
          v = s_best_v[w];
          r = s_best_r[w];
          c = s_best_c[w];
        }
      }
      Cand cd;
      cd.v = v;
      cd.r = r;
      cd.c = c;
      cd.pad = 0;
      cand[(size_t)(it & 1) * gridDim.x + blockIdx.x] = cd;
      __threadfence();
    }
    grid.sync();
    // 2. global pick: every CTA scans all candidates (first warp)
    if (wid == 0) {
      float v = -INFINITY;
      int r = 0x7fffffff, c = 0;
      for (int i = lane; i < (int)gridDim.x; i += 32) {
        const int4 raw = __ldcg(reinterpret_cast<const int4*>(&cand[(size_t)(it & 1) * gridDim.x + i]));
        Cand cd;
        cd.v = __int_as_float(raw.x);
        cd.r = raw.y;
        cd.c = raw.z;
        if (cd.r == 0x7fffffff) continue;
        if (r == 0x7fffffff || cd.v > v || (cd.v == v && cd.r < r)) {
          v = cd.v;
          r = cd.r;
          c = cd.c;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int orr = __shfl_xor_sync(0xffffffffu, r, o);
        int oc = __shfl_xor_sync(0xffffffffu, c, o);
        if (orr != 0x7fffffff && (r == 0x7fffffff || ov > v || (ov == v && orr < r))) {
          v = ov;
          r = orr;
          c = oc;
        }
      }
      if (lane == 0) {
        s_pick_r = r;
        s_pick_c = c;
      }
    }
    __syncthreads();
    const int pr = s_pick_r, pc = s_pick_c;
    if (blockIdx.x == 0 && threadIdx.x == 0) result[pr] = pc;
    // 3. preds[:, pc] = 0 ; preds[pr, :] = 0 ; rescale
    for (int r = wid; r < my_rows; r += nwarps) {
      float* row = rows + (size_t)r * C;
      if (row0 + r == pr) {
        for (int c = lane; c < C; c += 32) row[c] = 0.f;
      } else if (lane == 0) {
        row[pc] = 0.f;
      }
      __syncwarp();
      rescale_row(row, plan, leaf_sums, lane);
    }
    __syncthreads();
  }
}

}  // namespace rxb

extern "C" {

int rxb_tta_softmax_avg_mask(const float* logits, int V, int N, int C, const int32_t* plate,
                             const int32_t* group_col, float* probs, rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(logits && probs, "rxb_tta_softmax_avg_mask: null pointer");
  RXB_CHECK_ARG(V >= 1 && N >= 0 && C >= 1, "rxb_tta_softmax_avg_mask: bad sizes");
  RXB_CHECK_ARG((plate == nullptr) == (group_col == nullptr), "rxb_tta_softmax_avg_mask: plate/group_col mismatch");
  if (N == 0) return RXB_OK;
  int rc = rxb_check_device();
  if (rc) return rc;
  PairwisePlan plan;
  if (make_plan(plan, C)) return set_error(RXB_ERR_INVALID, "row length %d unsupported", C);
  size_t smem = (size_t)(C + kMaxLeaves) * sizeof(float);
  RXB_PROF(as_stream(stream), PROF_TTA);
  tta_softmax_avg_mask_kernel<<<N, kTtaThreads, smem, as_stream(stream)>>>(logits, V, N, plate, group_col,
                                                                             probs, plan, 1);
  RXB_LAUNCH_OK();
  return RXB_OK;
}

int rxb_mask_rescale(float* preds, int N, int C, const int32_t* plate, const int32_t* group_col,
                     rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(preds, "rxb_mask_rescale: null pointer");
  RXB_CHECK_ARG(N >= 0 && C >= 1, "rxb_mask_rescale: bad sizes");
  RXB_CHECK_ARG((plate == nullptr) == (group_col == nullptr), "rxb_mask_rescale: plate/group_col mismatch");
  if (N == 0) return RXB_OK;
  int rc = rxb_check_device();
  if (rc) return rc;
  PairwisePlan plan;
  if (make_plan(plan, C)) return set_error(RXB_ERR_INVALID, "row length %d unsupported", C);
  size_t smem = (size_t)(C + kMaxLeaves) * sizeof(float);
  RXB_PROF(as_stream(stream), PROF_TTA);
  tta_softmax_avg_mask_kernel<<<N, kTtaThreads, smem, as_stream(stream)>>>(nullptr, 1, N, plate, group_col,
                                                                             preds, plan, 0);
  RXB_LAUNCH_OK();
  return RXB_OK;
}

size_t rxb_greedy_assign_workspace_bytes(int N, int C) {
  (void)N;
  (void)C;
  return 2 * 1024 * sizeof(rxb::Cand);  // two candidate arrays, up to 1024 CTAs
}

int rxb_greedy_assign(const float* preds, int N, int C, int32_t* result, void* workspace,
                      rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(preds && result && workspace, "rxb_greedy_assign: null pointer");
  RXB_CHECK_ARG(N >= 0 && C >= 1, "rxb_greedy_assign: bad sizes");
  if (N == 0) return RXB_OK;
  int rc = rxb_check_device();
  if (rc) return rc;
  PairwisePlan plan;
  if (make_plan(plan, C)) return set_error(RXB_ERR_INVALID, "row length %d unsupported", C);
  int grid = num_sms();
  if (grid > 1024) grid = 1024;
  if (grid > N) grid = N;
  int rows_per_cta = ceil_div(N, grid);
  grid = ceil_div(N, rows_per_cta);
  size_t smem = ((size_t)rows_per_cta * C + (kGreedyThreads / 32) * kMaxLeaves) * sizeof(float);
  if (smem > 220 * 1024)
    return set_error(RXB_ERR_UNSUPPORTED, "rxb_greedy_assign: N=%d C=%d needs %zu B of shared memory per CTA", N,
                     C, smem);
  RXB_CUDA(cudaFuncSetAttribute(greedy_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  Cand* cand = reinterpret_cast<Cand*>(workspace);
  void* args[] = {(void*)&preds, (void*)&N, (void*)&rows_per_cta, (void*)&result, (void*)&cand, (void*)&plan};
  RXB_PROF(as_stream(stream), PROF_TTA);
  RXB_CUDA(cudaLaunchCooperativeKernel((void*)greedy_assign_kernel, dim3(grid), dim3(kGreedyThreads), args, smem,
                                       as_stream(stream)));
  ++g_launches;
  return RXB_OK;
}

}  // extern "C"
