// Device code of the sampler adaptor: Philox, stretch-move propose / accept (whole ensemble and per-device
// parts), slice broadcast, and the persistent cooperative ensemble kernel K5.  Included by sterne_b200.cu.
namespace {

// ---------------------------------------------------------------------------------
// Stretch-move sampler steps (sampler adaptor)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// uniform in (0, 1) from 53 random bits
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {
    const uint64_t k = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)k + 0.5) * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(kThreads)
stn_stretch_propose_kernel(const double* __restrict__ x, const int64_t W, const int P, const int64_t ld,
                           const int half, const double a, const uint64_t seed, const uint64_t step,
                           double* __restrict__ y, double* __restrict__ z_out) {
    const int64_t H = W / 2;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= H) return;
    uint32_t r[4];
    philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ (half ? 0x80000000u : 0u),
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u = u01(r[0], r[1]);
    const double t = (a - 1.0) * u + 1.0;
    const double z = t * t / a;                                       // g(z) ~ 1/sqrt(z) on [1/a, a]
    const int64_t j = (int64_t)(u01(r[2], r[3]) * (double)H);         // partner in the complementary half
    const double* xi = x + (half ? H + k : k) * ld;
    const double* xj = x + (half ? j : H + j) * ld;
    for (int p = 0; p < P; ++p) y[k * P + p] = fma(z, xi[p] - xj[p], xj[p]);
    z_out[k] = z;
}

__global__ void __launch_bounds__(kThreads)
stn_stretch_accept_kernel(double* __restrict__ x, double* __restrict__ lp, const int64_t W, const int P,
                          const int64_t ld, const int half, const uint64_t seed, const uint64_t step,
                          const double* __restrict__ y, const double* __restrict__ z, const double* __restrict__ lpy,
                          unsigned long long* __restrict__ n_accept) {
    const int64_t H = W / 2;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool acc = false;
    if (k < H) {
        uint32_t r[4];
        philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)step,
                      (uint32_t)(step >> 32) ^ (half ? 0xC0000000u : 0x40000000u), (uint32_t)seed,
                      (uint32_t)(seed >> 32), r);
        const int64_t i = half ? H + k : k;
        const double lnq = (double)(P - 1) * log(z[k]) + lpy[k] - lp[i];
        acc = (lpy[k] > -CUDART_INF) && (lpy[k] == lpy[k]) && (log(u01(r[0], r[1])) < lnq);
        if (acc) {
            for (int p = 0; p < P; ++p) x[i * ld + p] = y[k * P + p];
            lp[i] = lpy[k];
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(n_accept, (unsigned long long)__popc(ballot));
}

// Walker-sharded forms of the two stretch kernels (several GPUs, one ensemble): this device proposes for
// the proposals [k0, k0 + cnt) of the half-step from ITS copy of the walkers, and the accept kernel
// stores every accepted walker into the copies of ALL devices (peer-mapped pointers: the exchange of
// the updated half over NVLink is the kernel's own stores).  Philox counters are global walker
// indices, so the chain does not depend on how many devices share the ensemble.
struct WalkerCopies {
    double* x[STN_MAX_OUTS];
    double* lp[STN_MAX_OUTS];
    int n;
};

__global__ void __launch_bounds__(kThreads)
stn_stretch_propose_part_kernel(const double* __restrict__ x, const int64_t W, const int P, const int64_t ld,
                                const int half, const double a, const uint64_t seed, const uint64_t step,
                                const int64_t k0, const int64_t cnt, double* __restrict__ y, double* __restrict__ z_out) {
    const int64_t H = W / 2;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= cnt) return;
    const int64_t k = k0 + q;
    uint32_t r[4];
    philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ (half ? 0x80000000u : 0u),
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u = u01(r[0], r[1]);
    const double t = (a - 1.0) * u + 1.0;
    const double z = t * t / a;
    const int64_t j = (int64_t)(u01(r[2], r[3]) * (double)H);
    const double* xi = x + (half ? H + k : k) * ld;
    const double* xj = x + (half ? j : H + j) * ld;
    for (int p = 0; p < P; ++p) y[q * P + p] = fma(z, xi[p] - xj[p], xj[p]);
    z_out[q] = z;
}

__global__ void __launch_bounds__(kThreads)
stn_stretch_accept_part_kernel(const WalkerCopies C, const int64_t W, const int P, const int64_t ld, const int half,
                               const uint64_t seed, const uint64_t step, const int64_t k0, const int64_t cnt,
                               const double* __restrict__ y, const double* __restrict__ z,
                               const double* __restrict__ lpy, unsigned long long* __restrict__ n_accept) {
    const int64_t H = W / 2;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool acc = false;
    if (q < cnt) {
        const int64_t k = k0 + q;
        uint32_t r[4];
        philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)step,
                      (uint32_t)(step >> 32) ^ (half ? 0xC0000000u : 0x40000000u), (uint32_t)seed,
                      (uint32_t)(seed >> 32), r);
        const int64_t i = half ? H + k : k;
        const double lnq = (double)(P - 1) * log(z[q]) + lpy[q] - C.lp[0][i];
        acc = (lpy[q] > -CUDART_INF) && (lpy[q] == lpy[q]) && (log(u01(r[0], r[1])) < lnq);
        if (acc) {
            for (int c = 0; c < C.n; ++c) {
                for (int p = 0; p < P; ++p) C.x[c][i * ld + p] = y[q * P + p];
                C.lp[c][i] = lpy[q];
            }
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(n_accept, (unsigned long long)__popc(ballot));
}

// After a half-step, this device's slice of the updated half -- rows [row0, row0 + cnt) of x and of lp, a
// contiguous block -- goes to every other device's copy with coalesced 16-byte stores over NVLink
// (peer-mapped pointers; scattered 8-byte peer stores from the accept kernel itself were 3x slower).
__global__ void __launch_bounds__(kThreads)
stn_broadcast_rows_kernel(const WalkerCopies C, const int64_t row0, const int64_t cnt, const int64_t ld) {
    const int64_t nx = cnt * ld, stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double* sx = C.x[0] + row0 * ld;
    const double* sl = C.lp[0] + row0;
    if ((((uintptr_t)sx) & 15) == 0 && (nx & 1) == 0) {
        const double2* s2 = reinterpret_cast<const double2*>(sx);
        for (int64_t e = t0; e < nx / 2; e += stride) {
            const double2 v = s2[e];
            for (int c = 1; c < C.n; ++c) reinterpret_cast<double2*>(C.x[c] + row0 * ld)[e] = v;
        }
    } else {
        for (int64_t e = t0; e < nx; e += stride) {
            const double v = sx[e];
            for (int c = 1; c < C.n; ++c) C.x[c][row0 * ld + e] = v;
        }
    }
    for (int64_t e = t0; e < cnt; e += stride) {
        const double v = sl[e];
        for (int c = 1; c < C.n; ++c) C.lp[c][row0 + e] = v;
    }
}

// ---------------------------------------------------------------------------------
// Fused ensemble sampler: the WHOLE stretch-move loop in one persistent cooperative kernel.
// Per half-step every thread proposes for its walkers, evaluates log prior + log L of the
// proposal inline (the same per-source building blocks as K1, rows read straight from the
// L1/L2-resident epoch tables: these are the small real-data problems, ~10 epochs per source,
// where six launches per step cost more than the arithmetic) and accepts / rejects; the two
// half-steps of a step are separated by one grid-wide barrier.  Same Philox counters and the
// same arithmetic (plan lanes = 1, splits = 1) as the three-launch path, hence the same chain.
// ---------------------------------------------------------------------------------
#define STN_SAMPLER_MAX_P 64

template <bool COH>
__device__ double logpost_one(const ProbDev& P, const double* theta_row, int64_t ld) {
    double logl[1] = {0.0};
    const int64_t nidx[1] = {0};
    SrcState<1> st;
    st.mask = kChunk - 1;
    st.nvar = 0;
    for (int c = 0; c < P.n_chunks; ++c) {
        const ChunkDev ck = P.chunks[c];
        const SrcDev& S = P.src[ck.source];
        if (ck.flags & kFirst) src_begin<1, COH>(P, S, theta_row, nidx, ld, STN_LAYOUT_AOS, st);
        src_chunk<1, 1>(S, ck.rows, ck.n_epochs, 0, st);
        if (ck.flags & kLast) src_fold<1, 1>(P, S, true, st, logl);
    }
    double v = logl[0];
    if (P.n_a1dot > 0 || P.n_sini > 0) v += constraint_terms<COH>(P, theta_row, 0, ld, STN_LAYOUT_AOS);
    if (P.n_prior > 0) {
        const double lp = prior_terms<COH>(P, theta_row, 0, ld, STN_LAYOUT_AOS);
        v = (lp == -CUDART_INF) ? lp : v + lp;
    }
    return v;
}

// The same log-posterior with the SOURCES of one walker spread over LS lanes of a warp (lane `sub` of the group
// takes sources sub, sub + LS, ...): the per-source prologue, epoch loop and fold are independent, so a joint fit
// of S sources takes about the time of its longest source instead of the sum.  Every lane returns the full
// value: the two terms of each source are fetched from their owner with shuffles and added in source order,
// exactly as the one-lane form (and K1) adds them -- same bits.
constexpr int kSamplerMaxSourcesPerLane = 8;

template <bool COH>
__device__ double logpost_lanes(const ProbDev& P, const double* theta_row, int64_t ld, int sub, int LS) {
    double ta[kSamplerMaxSourcesPerLane], tb[kSamplerMaxSourcesPerLane];
#pragma unroll
    for (int j = 0; j < kSamplerMaxSourcesPerLane; ++j) ta[j] = tb[j] = 0.0;
    const int64_t nidx[1] = {0};
    SrcState<1> st;
    st.mask = kChunk - 1;
    st.nvar = 0;
    // lane `sub` walks ITS j-th source while the other lanes of the group walk theirs: the same code at the same
    // time on different sources (a loop over the common chunk list would serialise the lanes instead)
    for (int j = 0; j * LS < P.n_sources; ++j) {
        const int s = j * LS + sub;
        if (s >= P.n_sources) continue;
        const SrcDev& S = P.src[s];
        src_begin<1, COH, false>(P, S, theta_row, nidx, ld, STN_LAYOUT_AOS, st);
        for (int c = P.first_chunk[s]; c < P.first_chunk[s + 1]; ++c) {
            const ChunkDev ck = P.chunks[c];
            src_chunk<1, 1>(S, ck.rows, ck.n_epochs, 0, st);
        }
        double a[1], b[1];
        src_fold_terms<1, 1>(P, S, true, st, a, b);
        ta[j] = a[0];
        tb[j] = b[0];
    }
    __syncwarp();
    const int base = (threadIdx.x & 31) & ~(LS - 1);
    double v = 0.0;
    for (int j = 0; j * LS < P.n_sources; ++j)
        for (int o = 0; o < LS && j * LS + o < P.n_sources; ++o) {
            const double va = __shfl_sync(0xffffffffu, ta[j], base + o);
            const double vb = __shfl_sync(0xffffffffu, tb[j], base + o);
            v += va;
            v += vb;
        }
    if (P.n_a1dot > 0 || P.n_sini > 0) v += constraint_terms<COH>(P, theta_row, 0, ld, STN_LAYOUT_AOS);
    if (P.n_prior > 0) {
        const double lp = prior_terms<COH>(P, theta_row, 0, ld, STN_LAYOUT_AOS);
        v = (lp == -CUDART_INF) ? lp : v + lp;
    }
    return v;
}

// Bytes of shared memory sampler_stage_problem needs (host and device agree through this one function).
__host__ __device__ inline size_t sampler_align(size_t v) { return (v + 15) & ~(size_t)15; }

__device__ __forceinline__ void block_copy(void* dst, const void* src, size_t bytes) {
    // bytes is a multiple of 4 for every array copied here
    uint32_t* d = static_cast<uint32_t*>(dst);
    const uint32_t* q = static_cast<const uint32_t*>(src);
    for (size_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) d[i] = q[i];
}

// Copies everything ProbDev points to into shared memory (every CTA its own copy) and returns a ProbDev
// whose pointers refer to the copies.  Layout: SrcDev[S] | ChunkDev[C] | fast tables | constraint and prior arrays.
__device__ ProbDev sampler_stage_problem(const ProbDev& P, unsigned char* smem) {
    ProbDev Q = P;
    size_t off = 0;
    SrcDev* src = reinterpret_cast<SrcDev*>(smem + off);
    off += sampler_align(sizeof(SrcDev) * P.n_sources);
    ChunkDev* chunks = reinterpret_cast<ChunkDev*>(smem + off);
    off += sampler_align(sizeof(ChunkDev) * P.n_chunks);
    block_copy(src, P.src, sizeof(SrcDev) * P.n_sources);
    block_copy(chunks, P.chunks, sizeof(ChunkDev) * P.n_chunks);
    __syncthreads();
    // tables: one after the other in source order
    size_t table_off = off;
    for (int s = 0; s < P.n_sources; ++s) {
        const size_t bytes = sizeof(double) * STN_FAST_ROW * (size_t)src[s].n_epochs;
        block_copy(smem + table_off, P.src[s].fast_rows, bytes);
        table_off += sampler_align(bytes);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        size_t o = off;
        for (int s = 0; s < P.n_sources; ++s) {
            const double* old_rows = src[s].fast_rows;
            const double* new_rows = reinterpret_cast<const double*>(smem + o);
            for (int c = 0; c < P.n_chunks; ++c)
                if (chunks[c].source == s) chunks[c].rows = new_rows + (chunks[c].rows - old_rows);
            src[s].fast_rows = new_rows;
            o += sampler_align(sizeof(double) * STN_FAST_ROW * (size_t)src[s].n_epochs);
        }
    }
    off = table_off;
    auto take = [&](const void* g, size_t bytes) -> unsigned char* {
        unsigned char* d = smem + off;
        block_copy(d, g, bytes);
        off += sampler_align(bytes);
        return d;
    };
    Q.src = src;
    Q.chunks = chunks;
    Q.first_chunk = reinterpret_cast<const int*>(take(P.first_chunk, sizeof(int) * (P.n_sources + 1)));
    Q.a1_src = reinterpret_cast<const int*>(take(P.a1_src, sizeof(int) * P.n_a1dot));
    Q.a1_mu = reinterpret_cast<const double*>(take(P.a1_mu, sizeof(double) * P.n_a1dot));
    Q.a1_sigma = reinterpret_cast<const double*>(take(P.a1_sigma, sizeof(double) * P.n_a1dot));
    Q.sini_col = reinterpret_cast<const int*>(take(P.sini_col, sizeof(int) * P.n_sini));
    Q.sini_mu = reinterpret_cast<const double*>(take(P.sini_mu, sizeof(double) * P.n_sini));
    Q.sini_sigma = reinterpret_cast<const double*>(take(P.sini_sigma, sizeof(double) * P.n_sini));
    Q.prior_kind = reinterpret_cast<const int*>(take(P.prior_kind, sizeof(int) * P.n_prior));
    Q.prior_a = reinterpret_cast<const double*>(take(P.prior_a, sizeof(double) * P.n_prior));
    Q.prior_b = reinterpret_cast<const double*>(take(P.prior_b, sizeof(double) * P.n_prior));
    Q.sinlim_col = reinterpret_cast<const int*>(take(P.sinlim_col, sizeof(int) * P.n_sinlim));
    Q.sinlim_lo = reinterpret_cast<const double*>(take(P.sinlim_lo, sizeof(double) * P.n_sinlim));
    Q.sinlim_hi = reinterpret_cast<const double*>(take(P.sinlim_hi, sizeof(double) * P.n_sinlim));
    __syncthreads();
    return Q;
}

constexpr int kSamplerThreads = 256;

// All CTAs of the launch meet here.  cluster_mode: the launch is ONE thread-block cluster (up to 16 CTAs
// = 4096 proposals per half-step) and this is the hardware cluster barrier (release / acquire: the walkers
// just written by other CTAs are visible afterwards); otherwise a cooperative launch and the grid barrier.
__device__ __forceinline__ void sampler_barrier(bool cluster_mode) {
    namespace cg = cooperative_groups;
    if (cluster_mode) {
        asm volatile("barrier.cluster.arrive.release;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    } else {
        cg::this_grid().sync();
    }
}

__global__ void __launch_bounds__(kSamplerThreads)
stn_ensemble_kernel(const ProbDev P0, double* x, double* lp, const int64_t W, const int64_t ld, const double a,
                    const uint64_t seed, const uint64_t step0, const int64_t n_steps, const int64_t thin,
                    double* chain, double* lnprob, unsigned long long* n_accept, const int cluster_mode,
                    const int LS) {
    // The whole problem description -- sources, chunk list, epoch tables, constraints, prior -- is copied
    // to shared memory once: the barrier between half-steps invalidates L1 (acquire), so anything left in
    // global memory would be re-fetched from L2 (~700 cycles per dependent load) in every half-step.
    extern __shared__ __align__(16) unsigned char sampler_smem[];
    const ProbDev Q = sampler_stage_problem(P0, sampler_smem);
    const ProbDev& P = Q;
    const int np = P.n_params;
    const int64_t H = W / 2;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t Hpad = (H + 31) / 32 * 32;             // whole warps walk the loop together
    unsigned long long accepted = 0;
    double y[STN_SAMPLER_MAX_P];

    // Everything of a proposal that does not depend on the half-step that is still being written: the random
    // numbers, this walker's own row (it belongs to the half that is NOT being updated) and its log-posterior.
    struct Draw {
        int64_t i, jrow;
        double z, logz, lnu, lp_i;
        bool valid;
    };
    auto prepare = [&](int64_t kk, uint64_t step, int half, Draw& d) {
        d.valid = kk < H;
        const int64_t k = d.valid ? kk : H - 1;
        d.i = half ? H + k : k;
        const double* xi = x + d.i * ld;
        for (int p = 0; p < np; ++p) y[p] = __ldcg(xi + p);
        d.lp_i = __ldcg(lp + d.i);
        uint32_t r[4];
        philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ (half ? 0x80000000u : 0u),
                      (uint32_t)seed, (uint32_t)(seed >> 32), r);
        const double u = u01(r[0], r[1]);
        const double tt = (a - 1.0) * u + 1.0;
        d.z = tt * tt / a;
        d.logz = log(d.z);
        d.jrow = (half ? 0 : H) + (int64_t)(u01(r[2], r[3]) * (double)H);
        philox4x32_10((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)step,
                      (uint32_t)(step >> 32) ^ (half ? 0xC0000000u : 0x40000000u), (uint32_t)seed, (uint32_t)(seed >> 32), r);
        d.lnu = log(u01(r[0], r[1]));
    };
    // The part that needs the partner's CURRENT position: proposal, log-posterior, accept / reject.
    auto finish = [&](const Draw& d, int sub) {
        const double* xj = x + d.jrow * ld;
        for (int p = 0; p < np; ++p) {
            const double vj = __ldcg(xj + p);
            y[p] = fma(d.z, y[p] - vj, vj);
        }
        const double lpy = LS > 1 ? logpost_lanes<true>(P, y, np, sub, LS) : logpost_one<true>(P, y, np);
        const double lnq = (double)(np - 1) * d.logz + lpy - d.lp_i;
        if (d.valid && sub == 0 && (lpy > -CUDART_INF) && (lpy == lpy) && (d.lnu < lnq)) {
            for (int p = 0; p < np; ++p) x[d.i * ld + p] = y[p];
            lp[d.i] = lpy;
            ++accepted;
        }
    };
    auto store_chain = [&](int64_t t) {
        const int64_t slot = (t + 1) / thin - 1;
        for (int64_t e = gtid; e < W * np; e += nthreads) {
            const int64_t w = e / np;
            chain[slot * W * np + e] = __ldcg(x + w * ld + (e - w * np));
        }
        if (lnprob != nullptr)
            for (int64_t w = gtid; w < W; w += nthreads) lnprob[slot * W + w] = __ldcg(lp + w);
    };

    if (cluster_mode != 0 && LS == 1 && Hpad <= nthreads) {
        // One proposal per thread and a hardware cluster barrier: the barrier is SPLIT.  After a thread has
        // stored its accepted walker it arrives, prepares its next proposal (two Philox draws, its own row from
        // L2, two logarithms) while the other CTAs finish, and only then waits -- the barrier's latency and the
        // L2 round trip hide behind work that is needed anyway.
        const bool active = gtid < Hpad;
        Draw d;
        if (active) prepare(gtid, step0, 0, d);
        for (int64_t t = 0; t < n_steps; ++t) {
            for (int half = 0; half < 2; ++half) {
                if (active) finish(d, 0);
                asm volatile("barrier.cluster.arrive.release;" ::: "memory");
                const bool more = !(t == n_steps - 1 && half == 1);
                if (active && more) prepare(gtid, step0 + (uint64_t)(half ? t + 1 : t), half ^ 1, d);
                asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
            }
            if (chain != nullptr && (t + 1) % thin == 0) {
                store_chain(t);
                sampler_barrier(true);              // half-step 0 of the next step overwrites rows that are being copied
            }
        }
    } else {
        // general form: a thread may walk several proposals; with LS > 1 a group of LS lanes shares one proposal
        // (every lane draws the same numbers and builds the same proposal, the sources are split, lane 0 stores)
        const int64_t slot = gtid / LS, n_slots = nthreads / LS;
        const int sub = (int)(gtid % LS);
        const int per_warp = 32 / LS;
        const int64_t Hgrp = (H + per_warp - 1) / per_warp * per_warp;      // whole warps walk the loop together
        for (int64_t t = 0; t < n_steps; ++t) {
            const uint64_t step = step0 + (uint64_t)t;
            for (int half = 0; half < 2; ++half) {
                for (int64_t kk = slot; kk < Hgrp; kk += n_slots) {
                    Draw d;
                    prepare(kk, step, half, d);
                    finish(d, sub);
                }
                sampler_barrier(cluster_mode != 0);        // the other half reads these walkers next
            }
            if (chain != nullptr && (t + 1) % thin == 0) {
                store_chain(t);
                sampler_barrier(cluster_mode != 0);        // half-step 0 of the next step overwrites rows that are being copied
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) accepted += __shfl_xor_sync(0xffffffffu, accepted, o);
    if ((threadIdx.x & 31) == 0 && accepted) atomicAdd(n_accept, accepted);
}

}  // namespace
