// C ABI, part 3: the stretch-move sampler on one device (single kernels and stn_ensemble_run).
// Included by sterne_b200.cu.
extern "C" {

int stn_stretch_propose(int device, const double* x_dev, int64_t n_walkers, int32_t n_params, int64_t ld,
                        int half, double a, uint64_t seed, uint64_t step, double* y_dev, double* z_dev,
                        void* stream) {
    if (!x_dev || !y_dev || !z_dev) return fail(STN_ERR_INVALID, "null buffer");
    if (n_walkers < 2 || (n_walkers & 1) || n_params <= 0 || ld < n_params || !(a > 1.0) || (half != 0 && half != 1))
        return fail(STN_ERR_INVALID, "stretch move needs an even number of walkers >= 2, ld >= P and a > 1");
    STN_CUDA(cudaSetDevice(device));
    const int64_t H = n_walkers / 2;
    (void)cudaGetLastError();
    stn_stretch_propose_kernel<<<(unsigned)((H + kThreads - 1) / kThreads), kThreads, 0,
                                 static_cast<cudaStream_t>(stream)>>>(x_dev, n_walkers, n_params, ld, half, a, seed,
                                                                      step, y_dev, z_dev);
    STN_CUDA(cudaGetLastError());
    return STN_OK;
}

int stn_stretch_accept(int device, double* x_dev, double* lp_dev, int64_t n_walkers, int32_t n_params,
                       int64_t ld, int half, uint64_t seed, uint64_t step, const double* y_dev,
                       const double* z_dev, const double* lpy_dev, int64_t* n_accept_dev, void* stream) {
    if (!x_dev || !lp_dev || !y_dev || !z_dev || !lpy_dev || !n_accept_dev) return fail(STN_ERR_INVALID, "null buffer");
    if (n_walkers < 2 || (n_walkers & 1) || n_params <= 0 || ld < n_params || (half != 0 && half != 1))
        return fail(STN_ERR_INVALID, "stretch move needs an even number of walkers >= 2 and ld >= P");
    STN_CUDA(cudaSetDevice(device));
    const int64_t H = n_walkers / 2;
    (void)cudaGetLastError();
    stn_stretch_accept_kernel<<<(unsigned)((H + kThreads - 1) / kThreads), kThreads, 0,
                                static_cast<cudaStream_t>(stream)>>>(
        x_dev, lp_dev, n_walkers, n_params, ld, half, seed, step, y_dev, z_dev, lpy_dev,
        reinterpret_cast<unsigned long long*>(n_accept_dev));
    STN_CUDA(cudaGetLastError());
    return STN_OK;
}

int stn_ensemble_run(StnCtx* ctx, double* x_dev, double* lp_dev, int64_t n_walkers, int64_t ld, double a,
                     uint64_t seed, uint64_t step0, int64_t n_steps, int64_t thin, double* chain_dev,
                     double* lnprob_dev, int64_t* n_accept_dev, int mode, int plan, int engine, void* stream) {
    if (!ctx || !x_dev || !lp_dev || !n_accept_dev) return fail(STN_ERR_INVALID, "null argument");
    if (!ctx->has_prior) return fail(STN_ERR_INVALID, "stn_ensemble_run needs stn_set_prior first");
    const int P = ctx->prob.n_params;
    if (n_walkers < 2 || (n_walkers & 1) || ld < P || !(a > 1.0) || n_steps < 0 || thin < 1)
        return fail(STN_ERR_INVALID, "stretch move needs an even number of walkers >= 2, ld >= P, a > 1, thin >= 1");
    if (mode != STN_MODE_FAST && mode != STN_MODE_STRICT) return fail(STN_ERR_INVALID, "bad mode %d", mode);
    if (n_steps == 0) return STN_OK;
    STN_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t H = n_walkers / 2;
    // engine 0: choose; 1: persistent cooperative kernel; 2: three launches per half-step
    int64_t evals = 0;
    for (int e : ctx->n_epochs) evals += e;
    const bool can_fuse = mode == STN_MODE_FAST && P <= STN_SAMPLER_MAX_P && (plan == 0 || plan == make_plan(1, 1));
    if (engine == 1 && !can_fuse)
        return fail(STN_ERR_INVALID, "the persistent sampler kernel needs mode fast, plan 0 or lanes=1|splits=1, P <= %d",
                    STN_SAMPLER_MAX_P);
    bool fuse = engine == 1 || (engine == 0 && can_fuse && evals <= 512);
    int grid = 0, lanes = 1;
    size_t smem = 0;
    if (fuse) {
        // shared-memory image of the problem (sampler_stage_problem's layout)
        const ProbDev& pp = ctx->prior_prob;
        smem = sampler_align(sizeof(SrcDev) * pp.n_sources) + sampler_align(sizeof(ChunkDev) * pp.n_chunks) +
               sampler_align(sizeof(int) * (pp.n_sources + 1));
        for (int e : ctx->n_epochs) smem += sampler_align(sizeof(double) * STN_FAST_ROW * (size_t)e);
        smem += 3 * sampler_align(8 * (size_t)pp.n_a1dot) + 3 * sampler_align(8 * (size_t)pp.n_sini) +
                3 * sampler_align(8 * (size_t)pp.n_prior) + 3 * sampler_align(8 * (size_t)pp.n_sinlim);
        if (smem > 200 * 1024) {
            if (engine == 1) return fail(STN_ERR_INVALID, "the problem (%zu bytes) does not fit the persistent kernel's shared memory", smem);
            fuse = false;
        }
    }
    if (fuse) {
        int per_sm = 0;
        STN_CUDA(cudaFuncSetAttribute(stn_ensemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        STN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stn_ensemble_kernel, kSamplerThreads, smem));
        // lanes per walker: the sources of a joint fit are spread over up to 8 lanes of a warp
        lanes = 1;
        if (STN_SAMPLER_SOURCE_LANES && ctx->prob.n_sources > 1) {
            while (lanes < 8 && lanes < ctx->prob.n_sources) lanes *= 2;
            if ((ctx->prob.n_sources + lanes - 1) / lanes > kSamplerMaxSourcesPerLane) lanes = 1;
        }
        const int64_t want = (H * lanes + kSamplerThreads - 1) / kSamplerThreads;
        const int64_t cap = (int64_t)per_sm * ctx->sm_count;
        if (cap < 1) {
            if (engine == 1) return fail(STN_ERR_CUDA, "the persistent sampler kernel does not fit on this device");
            fuse = false;
        }
        grid = (int)(want < cap ? want : cap);
    }
    if (fuse) {
        ProbDev prob = ctx->prior_prob;
        prob.out_chisq = 0;
        prob.n_extra_out = 0;
        unsigned long long* nacc = reinterpret_cast<unsigned long long*>(n_accept_dev);
        (void)cudaGetLastError();
        // up to 16 CTAs: one thread-block cluster, hardware barrier between half-steps
        if (grid <= 16 && STN_SAMPLER_CLUSTER) {
            bool ok = true;
            if (grid > 8)
                ok = cudaFuncSetAttribute(stn_ensemble_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
            if (ok) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)grid);
                cfg.blockDim = dim3(kSamplerThreads);
                cfg.dynamicSmemBytes = smem;
                cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = (unsigned)grid;
                at[0].val.clusterDim.y = 1;
                at[0].val.clusterDim.z = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                const cudaError_t e = cudaLaunchKernelEx(&cfg, stn_ensemble_kernel, prob, x_dev, lp_dev, n_walkers, ld, a, seed,
                                                         step0, n_steps, thin, chain_dev, lnprob_dev, nacc, 1, lanes);
                if (e == cudaSuccess) {
                    ctx->launches++;
                    return STN_OK;
                }
            }
            (void)cudaGetLastError();                     // cluster launch refused: fall back to the cooperative grid
        }
        int cluster_mode = 0;
        void* args[] = {&prob, &x_dev, &lp_dev, &n_walkers, &ld, &a, &seed, &step0, &n_steps, &thin, &chain_dev,
                        &lnprob_dev, &nacc, &cluster_mode, &lanes};
        STN_CUDA(cudaLaunchCooperativeKernel((const void*)stn_ensemble_kernel, dim3((unsigned)grid), dim3(kSamplerThreads), args, smem, st));
        ctx->launches++;
        return STN_OK;
    }
    // generic path (long series: the throughput kernel per half-step), launches issued from here
    // instead of from Python; proposals and their log-posteriors live in the context's scratch
    if (plan == 0) plan = pick_plan(ctx, H);
    const int64_t need = H * P + 2 * H;
    if (ctx->sampler_elems < need) {
        STN_CUDA(cudaStreamSynchronize(st));
        if (ctx->sampler_buf) cudaFree(ctx->sampler_buf);
    if (ctx->sampler_nacc) cudaFree(ctx->sampler_nacc);
    for (int i = 0; i < 2; ++i)
        if (ctx->sampler_ev[i]) cudaEventDestroy(ctx->sampler_ev[i]);
        ctx->sampler_buf = nullptr;
        ctx->sampler_elems = 0;
        STN_CUDA(cudaMalloc(&ctx->sampler_buf, (size_t)need * sizeof(double)));
        ctx->sampler_elems = need;
    }
    double* y = ctx->sampler_buf;
    double* z = y + H * P;
    double* lpy = z + H;
    for (int64_t t = 0; t < n_steps; ++t) {
        for (int half = 0; half < 2; ++half) {
            int rc = stn_stretch_propose(ctx->device, x_dev, n_walkers, P, ld, half, a, seed, step0 + (uint64_t)t, y, z, stream);
            if (rc) return rc;
            rc = launch_loglike(ctx, y, H, P, STN_LAYOUT_AOS, mode, plan, lpy, st, 1);
            if (rc) return rc;
            rc = stn_stretch_accept(ctx->device, x_dev, lp_dev, n_walkers, P, ld, half, seed, step0 + (uint64_t)t, y, z, lpy,
                                    n_accept_dev, stream);
            if (rc) return rc;
            ctx->launches += 2;
        }
        if (chain_dev && (t + 1) % thin == 0) {
            const int64_t slot = (t + 1) / thin - 1;
            STN_CUDA(cudaMemcpy2DAsync(chain_dev + slot * n_walkers * P, (size_t)P * sizeof(double), x_dev,
                                       (size_t)ld * sizeof(double), (size_t)P * sizeof(double), (size_t)n_walkers,
                                       cudaMemcpyDeviceToDevice, st));
            if (lnprob_dev)
                STN_CUDA(cudaMemcpyAsync(lnprob_dev + slot * n_walkers, lp_dev, (size_t)n_walkers * sizeof(double),
                                         cudaMemcpyDeviceToDevice, st));
        }
    }
    return STN_OK;
}

}  // extern "C"
