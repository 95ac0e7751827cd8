// reduce.cu -- K3: per-step cost accumulation + importance-sampling softmax + weighted control sum over the
// K trajectories as ONE fused single-pass reduction (north-star item (iii)), K4: shard combine + U update,
// and the U shift.  Replaces mppi_torch `_compute_rollout_costs` accumulation, `_exp_util`,
// `_update_distribution` (external dep mppi_torch@75e17e8; call site mppiisaac/planner/mppi_isaac.py:113;
// spec SURVEY.md 8(a) M5/M6 and 8(e)).
//
// K3 data flow (HBM-bound; every input byte is read from HBM exactly once):
//   cost[T][K], x[T*nu][K]  --128-bit coalesced loads-->  shared-memory tile of CK=64 samples
//   S_k   = sum_t gamma^t cost[t][k]  (+ lambda * sum_r g[r] x[r][k],  g = Sigma^-1 U,  SIMPLE mode)
//   beta_c = min_k S_k ; w_k = exp(-(S_k - beta_c)/lambda) ; eta_c = sum_k w_k
//   W_c[r] = sum_k w_k x[r][k]            (thread r, rotated float4 reads: bank-conflict free)
//   CTA running (beta, eta, W) merged online (log-sum-exp style rescale); CTAs grid-stride over chunks;
//   per-CTA partials -> global scratch; the LAST CTA (atomic ticket) folds them into `partial`.
// Algorithmic bytes per launch: 4*K*T*(nu+1) + 4*(T*nu+2)   (BASELINE.md section 3).
#include "common.cuh"

namespace {

constexpr int CK = 64;          // samples per tile
constexpr int NT = 256;         // threads per CTA
constexpr int RPT = 2;          // rows of W per thread (T*nu <= 512)

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// merge (b2, e2, w2) into running (b, e, w...) -- caller applies s_old / s_new to its W registers
__device__ __forceinline__ void merge_scales(float b_run, float b_new, float inv_lambda, float& b_out, float& s_old, float& s_new) {
    b_out = fminf(b_run, b_new);
    s_old = (b_run == INFINITY) ? 0.f : expf(-(b_run - b_out) * inv_lambda);
    s_new = (b_new == INFINITY) ? 0.f : expf(-(b_new - b_out) * inv_lambda);
}

__global__ void __launch_bounds__(NT)
reduce_kernel(const __grid_constant__ MppibParams p, int nu, const float* __restrict__ cost, const float* __restrict__ x,
              const float* __restrict__ U, float* __restrict__ scratch, unsigned int* __restrict__ ticket,
              float* __restrict__ partial) {
    extern __shared__ __align__(16) float smem[];
    const int K = p.K, T = p.T, NR = T * nu;
    const int tid = threadIdx.x;
    const float inv_lambda = 1.0f / p.lambda_;
    const bool simple = p.mode == MPPIB_MODE_SIMPLE;

    float* xs = smem;                    // [NR][CK]
    float* cs = xs + (size_t)NR * CK;    // [T][CK]
    float* g = cs + (size_t)T * CK;      // [NR]   lambda * Sigma^-1 U (SIMPLE)
    float* gp = g + NR;                  // [T]    gamma^t
    float* red = gp + T;                 // [4][CK]
    float* wk = red + 4 * CK;            // [CK]
    float* misc = wk + CK;               // [4]

    for (int r = tid; r < NR; r += NT) {
        float acc = 0.f;
        if (simple) {
            const int t = r / nu, i = r % nu;
            for (int j = 0; j < nu; ++j) acc += p.sigma_inv[i * nu + j] * U[t * nu + j];
            acc *= p.lambda_;
        }
        g[r] = acc;
    }
    for (int t = tid; t < T; t += NT) gp[t] = powf(p.gamma, (float)t);

    float b_run = INFINITY, e_run = 0.f, w_run[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) w_run[i] = 0.f;

    const int nchunks = (K + CK - 1) / CK;
    for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int k0 = chunk * CK;
        __syncthreads();   // previous tile fully consumed (also orders g/gp init on the first trip)
        // ---- stage the tile: 128-bit loads, k innermost => each row segment is 256 contiguous bytes
        const int nvec = (NR + T) * (CK / 4);
        for (int idx = tid; idx < nvec; idx += NT) {
            const int row = idx / (CK / 4), c4 = idx % (CK / 4);
            const int k = k0 + 4 * c4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* src = row < NR ? x + (size_t)row * K : cost + (size_t)(row - NR) * K;
            if (k + 3 < K) v = __ldg(reinterpret_cast<const float4*>(src + k));
            else if (k < K) { v.x = src[k]; if (k + 1 < K) v.y = src[k + 1]; if (k + 2 < K) v.z = src[k + 2]; }
            float* dst = row < NR ? xs + (size_t)row * CK : cs + (size_t)(row - NR) * CK;
            *reinterpret_cast<float4*>(dst + 4 * c4) = v;
        }
        __syncthreads();
        // ---- trajectory cost S_k: 4 partial sums per sample
        {
            const int kk = tid & (CK - 1), part = tid >> 6;
            float acc = 0.f;
            for (int t = part; t < T; t += 4) acc += gp[t] * cs[t * CK + kk];
            if (simple) for (int r = part; r < NR; r += 4) acc += g[r] * xs[r * CK + kk];
            red[part * CK + kk] = acc;
        }
        __syncthreads();
        if (tid < CK) {
            float S = red[tid] + red[CK + tid] + red[2 * CK + tid] + red[3 * CK + tid];
            const bool valid = (k0 + tid < K) && isfinite(S);
            S = valid ? S : INFINITY;
            float bmin = warp_min(S);
            if ((tid & 31) == 0) misc[tid >> 5] = bmin;
            red[tid] = S;
        }
        __syncthreads();
        const float b_c = fminf(misc[0], misc[1]);
        if (tid < CK) {
            const float S = red[tid];
            const float w = (S == INFINITY) ? 0.f : expf(-(S - b_c) * inv_lambda);
            wk[tid] = w;
            float es = warp_sum(w);
            if ((tid & 31) == 0) misc[2 + (tid >> 5)] = es;
        }
        __syncthreads();
        if (b_c != INFINITY) {
            const float e_c = misc[2] + misc[3];
            float b_out, s_old, s_new;
            merge_scales(b_run, b_c, inv_lambda, b_out, s_old, s_new);
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int r = tid + i * NT;
                if (r < NR) {
                    float acc = 0.f;
                    const float4* xr = reinterpret_cast<const float4*>(xs + (size_t)r * CK);
                    const float4* w4 = reinterpret_cast<const float4*>(wk);
#pragma unroll
                    for (int j = 0; j < CK / 4; ++j) {
                        const int jj = (j + r) & (CK / 4 - 1);   // rotation => conflict-free LDS.128
                        const float4 a = xr[jj], b = w4[jj];
                        acc += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
                    }
                    w_run[i] = w_run[i] * s_old + acc * s_new;
                }
            }
            e_run = e_run * s_old + e_c * s_new;
            b_run = b_out;
        }
    }
    // ---- per-CTA partial -> scratch ; last CTA folds
    const int P = 2 + NR;
    float* mine = scratch + (size_t)blockIdx.x * P;
    if (tid == 0) { mine[0] = b_run; mine[1] = e_run; }
#pragma unroll
    for (int i = 0; i < RPT; ++i) { const int r = tid + i * NT; if (r < NR) mine[2 + r] = w_run[i]; }
    __threadfence();
    __syncthreads();
    __shared__ unsigned int s_last;
    if (tid == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float b = INFINITY;
    for (int c = 0; c < (int)gridDim.x; ++c) b = fminf(b, __ldcg(scratch + (size_t)c * P));
    float e = 0.f, w[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) w[i] = 0.f;
    for (int c = 0; c < (int)gridDim.x; ++c) {
        const float* src = scratch + (size_t)c * P;
        const float bc = __ldcg(src);
        if (bc == INFINITY) continue;
        const float s = expf(-(bc - b) * inv_lambda);
        e += s * __ldcg(src + 1);
#pragma unroll
        for (int i = 0; i < RPT; ++i) { const int r = tid + i * NT; if (r < NR) w[i] += s * __ldcg(src + 2 + r); }
    }
    if (tid == 0) { partial[0] = b; partial[1] = e; *ticket = 0u; }
#pragma unroll
    for (int i = 0; i < RPT; ++i) { const int r = tid + i * NT; if (r < NR) partial[2 + r] = w[i]; }
}

// K4: combine G shard partials, U update, optional Savitzky-Golay (window 9, order 2, 'interp' edges), action out.
__global__ void __launch_bounds__(512)
finalize_kernel(const __grid_constant__ MppibParams p, int nu, const float* __restrict__ partials, int G,
                float* __restrict__ U, float* __restrict__ action_out, float* __restrict__ stats) {
    extern __shared__ float un[];   // [T*nu]
    const int T = p.T, NR = T * nu, P = 2 + NR;
    const float inv_lambda = 1.0f / p.lambda_;
    float b = INFINITY;
    for (int gidx = 0; gidx < G; ++gidx) if (partials[(size_t)gidx * P + 1] > 0.f) b = fminf(b, partials[(size_t)gidx * P]);
    float e = 0.f;
    for (int gidx = 0; gidx < G; ++gidx) {
        const float eg = partials[(size_t)gidx * P + 1];
        if (eg > 0.f) e += expf(-(partials[(size_t)gidx * P] - b) * inv_lambda) * eg;
    }
    for (int r = threadIdx.x; r < NR; r += blockDim.x) {
        float w = 0.f;
        for (int gidx = 0; gidx < G; ++gidx) {
            const float eg = partials[(size_t)gidx * P + 1];
            if (eg > 0.f) w += expf(-(partials[(size_t)gidx * P] - b) * inv_lambda) * partials[(size_t)gidx * P + 2 + r];
        }
        const float wm = e > 0.f ? w / e : (p.mode == MPPIB_MODE_SIMPLE ? 0.f : U[r]);   // no valid sample: keep U
        un[r] = p.mode == MPPIB_MODE_SIMPLE ? U[r] + wm : (1.0f - p.step_size_mean) * U[r] + p.step_size_mean * wm;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < NR; r += blockDim.x) {
        float out = un[r];
        if (p.filter_u) {
            const int t = r / nu, j = r % nu;
            const float mid[9] = {-21.f, 14.f, 39.f, 54.f, 59.f, 54.f, 39.f, 14.f, -21.f};
            const float edge[4][9] = {{763.f, 441.f, 189.f, 7.f, -105.f, -147.f, -119.f, -21.f, 147.f},
                                      {441.f, 322.f, 220.5f, 136.5f, 70.f, 21.f, -10.5f, -24.5f, -21.f},
                                      {189.f, 220.5f, 232.f, 223.5f, 195.f, 146.5f, 78.f, -10.5f, -119.f},
                                      {7.f, 136.5f, 223.5f, 268.f, 270.f, 229.5f, 146.5f, 21.f, -147.f}};
            float s = 0.f;
            if (t < 4) { for (int i = 0; i < 9; ++i) s += edge[t][i] * un[i * nu + j]; s *= (1.0f / 1155.0f); }
            else if (t >= T - 4) { const int ee = T - 1 - t; for (int i = 0; i < 9; ++i) s += edge[ee][i] * un[(T - 1 - i) * nu + j]; s *= (1.0f / 1155.0f); }
            else { for (int i = 0; i < 9; ++i) s += mid[i] * un[(t - 4 + i) * nu + j]; s *= (1.0f / 231.0f); }
            out = fminf(fmaxf(s, p.u_min[j]), p.u_max[j]);   // smoothing may overshoot the bounds at the edges
        }
        U[r] = out;
        if (r < nu) action_out[r] = out;
    }
    if (threadIdx.x == 0 && stats) { stats[0] = b; stats[1] = e; }
}

__global__ void shift_kernel(const __grid_constant__ MppibParams p, int nu, float* __restrict__ U, uint32_t* __restrict__ plan_ctr) {
    extern __shared__ float tmp[];
    const int NR = p.T * nu;
    for (int r = threadIdx.x; r < NR; r += blockDim.x) tmp[r] = r + nu < NR ? U[r + nu] : p.u_init[r % nu];
    __syncthreads();
    for (int r = threadIdx.x; r < NR; r += blockDim.x) U[r] = tmp[r];
    if (threadIdx.x == 0 && plan_ctr) *plan_ctr += 1u;
}

size_t reduce_smem_bytes(int T, int nu) {
    const int NR = T * nu;
    return sizeof(float) * ((size_t)NR * CK + (size_t)T * CK + NR + T + 4 * CK + CK + 4);
}

}  // namespace

int reduce_grid_size(const MppibContext* c) {
    const int nchunks = (c->params.K + CK - 1) / CK;
    int per_sm = (int)(220 * 1024 / reduce_smem_bytes(c->params.T, c->model.nu));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    const int cap = c->num_sms * per_sm;
    return nchunks < cap ? nchunks : cap;
}

int launch_reduce(MppibContext* c, const float* cost, const float* x, const float* U, float* partial, cudaStream_t s) {
    const int T = c->params.T, nu = c->model.nu;
    MPPIB_REQUIRE(T * nu <= RPT * NT, "mppib_reduce: T*nu = %d exceeds %d", T * nu, RPT * NT);
    const size_t smem = reduce_smem_bytes(T, nu);
    MPPIB_REQUIRE(smem <= 227 * 1024, "mppib_reduce: tile of %zu bytes exceeds shared memory", smem);
    static bool attr_set = false;
    if (!attr_set) {
        MPPIB_CHECK_CUDA(cudaFuncSetAttribute(reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int grid = reduce_grid_size(c);
    MPPIB_REQUIRE(grid <= c->reduce_max_ctas, "mppib_reduce: scratch too small");
    reduce_kernel<<<grid, NT, smem, s>>>(c->params, nu, cost, x, U, c->reduce_scratch, c->reduce_ticket, partial);
    MPPIB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_finalize(MppibContext* c, const float* partials, int G, float* U, float* action_out, float* stats, cudaStream_t s) {
    const int NR = c->params.T * c->model.nu;
    finalize_kernel<<<1, 512, NR * sizeof(float), s>>>(c->params, c->model.nu, partials, G, U, action_out, stats);
    MPPIB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_shift(MppibContext* c, float* U, uint32_t* plan_ctr, cudaStream_t s) {
    const int NR = c->params.T * c->model.nu;
    shift_kernel<<<1, 256, NR * sizeof(float), s>>>(c->params, c->model.nu, U, plan_ctr);
    MPPIB_CHECK_CUDA(cudaGetLastError());
    return 0;
}
