// C ABI, part 2: several devices driven by one process (host batches, device-resident slices, one walker
// ensemble shared by all devices).  Included by sterne_b200.cu.
extern "C" {

// ---------------------------------------------------------------------------------
// Several devices as one (single process, e.g. one sampler): contiguous slices of the batch, one
// per device, every device holding the full epoch tables.  No collective: for host batches each
// device DMAs its slice of the result straight into the caller's array; for device-resident
// slices the results are gathered with peer copies (NVLink) into one buffer.
// ---------------------------------------------------------------------------------
int stn_multi_create(const StnProblem* pb, const int* devices, int n_devices, StnMulti** out) {
    if (!pb || !devices || !out) return fail(STN_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n_devices <= 0 || n_devices > 64) return fail(STN_ERR_INVALID, "bad device count %d", n_devices);
    StnMulti* m = new (std::nothrow) StnMulti();
    if (!m) return fail(STN_ERR_NOMEM, "out of host memory");
    for (int g = 0; g < n_devices; ++g) {
        StnCtx* c = nullptr;
        const int rc = stn_create(pb, devices[g], &c);
        if (rc) {
            stn_multi_destroy(m);
            return rc;
        }
        m->ctx.push_back(c);
    }
    // peer access for stn_multi_loglike_dev: a kernel then stores its results straight into the
    // gathering device's buffer over NVLink (an already-enabled pair is not an error)
    m->peer_ok.assign((size_t)n_devices * n_devices, 0);
    for (int a = 0; a < n_devices; ++a)
        for (int b = 0; b < n_devices; ++b) {
            if (devices[a] == devices[b]) {
                m->peer_ok[(size_t)a * n_devices + b] = 1;
                continue;
            }
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) == cudaSuccess && can) {
                cudaSetDevice(devices[a]);
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
                if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) m->peer_ok[(size_t)a * n_devices + b] = 1;
            }
            (void)cudaGetLastError();
        }
    *out = m;
    return STN_OK;
}

int stn_multi_destroy(StnMulti* m) {
    if (!m) return STN_OK;
    for (StnCtx* c : m->ctx) stn_destroy(c);
    delete m;
    return STN_OK;
}

int stn_multi_size(StnMulti* m) { return m ? (int)m->ctx.size() : 0; }

StnCtx* stn_multi_ctx(StnMulti* m, int i) {
    if (!m || i < 0 || i >= (int)m->ctx.size()) return nullptr;
    return m->ctx[i];
}

int64_t stn_multi_launch_count(StnMulti* m) {
    int64_t t = 0;
    if (m) for (StnCtx* c : m->ctx) t += c->launches;
    return t;
}

int stn_multi_slice(int64_t n, int n_parts, int part, int64_t* lo, int64_t* hi) {
    if (n < 0 || n_parts <= 0 || part < 0 || part >= n_parts || !lo || !hi) return fail(STN_ERR_INVALID, "bad slice request");
    const int64_t base = n / n_parts, extra = n % n_parts;
    *lo = part * base + (part < extra ? part : extra);
    *hi = *lo + base + (part < extra ? 1 : 0);
    return STN_OK;
}

int stn_multi_loglike_host(StnMulti* m, const double* theta_host, int64_t n, int64_t ld, int layout,
                           int mode, int plan, double* out_host) {
    if (!m || m->ctx.empty()) return fail(STN_ERR_INVALID, "null multi-context");
    int rc = check_call(m->ctx[0], theta_host, n, ld, layout, mode, out_host);
    if (rc) return rc;
    if (n == 0) return STN_OK;
    // the plan fixes the summation order: chosen from the WHOLE batch so that the bits do not depend on
    // how many devices share it (R, the samples per thread, never changes a bit)
    if (mode == STN_MODE_FAST && plan == 0) plan = pick_plan(m->ctx[0], n);
    // devices used: a slice below 2^15 rows does not fill one B200
    int G = (int)m->ctx.size();
    const int64_t want = n >> 15;
    if (want < G) G = want < 1 ? 1 : (int)want;
    if (G == 1) {
        STN_CUDA(cudaSetDevice(m->ctx[0]->device));
        return run_host(m->ctx[0], theta_host, n, ld, layout, mode, plan, out_host);
    }
    std::vector<int> rcs(G, STN_OK);
    std::vector<std::string> errs(G);
    auto work = [&](int g) {
        int64_t lo = 0, hi = 0;
        stn_multi_slice(n, G, g, &lo, &hi);
        StnCtx* c = m->ctx[g];
        if (c->host_share < G) c->host_share = G;       // G devices copy from this host at once
        if (cudaSetDevice(c->device) != cudaSuccess) {
            rcs[g] = STN_ERR_CUDA;
            errs[g] = "cudaSetDevice failed";
            return;
        }
        const double* th = layout == STN_LAYOUT_AOS ? theta_host + lo * ld : theta_host + lo;
        rcs[g] = run_host(c, th, hi - lo, ld, layout, mode, plan, out_host + lo);
        if (rcs[g]) errs[g] = g_err;                    // g_err is thread-local
    };
    std::vector<std::thread> pool;
    pool.reserve(G - 1);
    for (int g = 1; g < G; ++g) pool.emplace_back(work, g);
    work(0);
    for (std::thread& t : pool) t.join();
    cudaSetDevice(m->ctx[0]->device);
    for (int g = 0; g < G; ++g)
        if (rcs[g]) return fail(rcs[g], "device slice %d: %s", g, errs[g].c_str());
    return STN_OK;
}

int stn_multi_loglike_dev(StnMulti* m, const double* const* theta_dev, const int64_t* counts, int64_t ld,
                          int layout, int mode, int plan, double* out_dev, int out_device) {
    if (!m || m->ctx.empty() || !theta_dev || !counts) return fail(STN_ERR_INVALID, "null argument");
    const int G = (int)m->ctx.size();
    int64_t n = 0;
    for (int g = 0; g < G; ++g) {
        if (counts[g] < 0) return fail(STN_ERR_INVALID, "negative slice size");
        if (counts[g] > 0 && !theta_dev[g]) return fail(STN_ERR_INVALID, "null slice");
        n += counts[g];
    }
    if (n == 0) return STN_OK;
    if (!out_dev) return fail(STN_ERR_INVALID, "null output");
    if (layout != STN_LAYOUT_AOS && layout != STN_LAYOUT_SOA) return fail(STN_ERR_INVALID, "bad layout %d", layout);
    if (mode != STN_MODE_FAST && mode != STN_MODE_STRICT) return fail(STN_ERR_INVALID, "bad mode %d", mode);
    if (mode == STN_MODE_FAST && plan == 0) plan = pick_plan(m->ctx[0], n);      // from the whole batch
    int rc = STN_OK;
    int64_t off = 0;
    for (int g = 0; g < G && !rc; ++g) {
        StnCtx* c = m->ctx[g];
        const int64_t cnt = counts[g];
        if (cnt == 0) continue;
        if ((rc = check_call(c, theta_dev[g], cnt, ld, layout, mode, out_dev))) break;
        STN_CUDA(cudaSetDevice(c->device));
        if ((rc = host_pipe_init(c))) break;
        cudaStream_t st = c->streams[0];
        bool direct = c->device == out_device;
        for (int b = 0; b < G && !direct; ++b)
            direct = m->ctx[b]->device == out_device && m->peer_ok[(size_t)g * G + b];
        if (direct) {
            // out_dev is local or peer-mapped: the kernel's own stores are the gather (8 bytes per vector over NVLink)
            rc = launch_loglike(c, theta_dev[g], cnt, ld, layout, mode, plan, out_dev + off, st, 0, 1);
        } else {
            if (c->gather_elems < cnt) {
                STN_CUDA(cudaStreamSynchronize(st));
                if (c->gather_buf) cudaFree(c->gather_buf);
                c->gather_buf = nullptr;
                c->gather_elems = 0;
                STN_CUDA(cudaMalloc(&c->gather_buf, (size_t)cnt * sizeof(double)));
                c->gather_elems = cnt;
            }
            rc = launch_loglike(c, theta_dev[g], cnt, ld, layout, mode, plan, c->gather_buf, st, 0, 1);
            if (!rc) {
                const cudaError_t e = cudaMemcpyPeerAsync(out_dev + off, out_device, c->gather_buf, c->device,
                                                          (size_t)cnt * sizeof(double), st);
                if (e != cudaSuccess) rc = fail(STN_ERR_CUDA, "peer copy %d -> %d failed: %s", c->device, out_device,
                                                cudaGetErrorString(e));
            }
        }
        off += cnt;
    }
    // synchronous on return (also after an error: nothing may still be running)
    for (int g = 0; g < G; ++g) {
        StnCtx* c = m->ctx[g];
        if (c->streams[0]) {
            cudaSetDevice(c->device);
            const cudaError_t e = cudaStreamSynchronize(c->streams[0]);
            if (e != cudaSuccess && !rc) rc = fail(STN_ERR_CUDA, "device %d: %s", c->device, cudaGetErrorString(e));
        }
    }
    cudaSetDevice(m->ctx[0]->device);
    return rc;
}

int stn_multi_set_prior(StnMulti* m, const StnPrior* prior) {
    if (!m || m->ctx.empty()) return fail(STN_ERR_INVALID, "null multi-context");
    for (StnCtx* c : m->ctx) {
        const int rc = stn_set_prior(c, prior);
        if (rc) return rc;
    }
    return STN_OK;
}

int stn_multi_ensemble_run(StnMulti* m, double* const* x_dev, double* const* lp_dev, int64_t n_walkers, int64_t ld,
                           double a, uint64_t seed, uint64_t step0, int64_t n_steps, int64_t thin, double* chain_dev,
                           double* lnprob_dev, int64_t* n_accept_host, int mode, int plan) {
    if (!m || m->ctx.empty() || !x_dev || !lp_dev) return fail(STN_ERR_INVALID, "null argument");
    const int G = (int)m->ctx.size();
    if (G > STN_MAX_OUTS) return fail(STN_ERR_INVALID, "at most %d devices can share one ensemble", STN_MAX_OUTS);
    const int P = m->ctx[0]->prob.n_params;
    if (n_walkers < 2 || (n_walkers & 1) || ld < P || !(a > 1.0) || n_steps < 0 || thin < 1)
        return fail(STN_ERR_INVALID, "stretch move needs an even number of walkers >= 2, ld >= P, a > 1, thin >= 1");
    if (mode != STN_MODE_FAST && mode != STN_MODE_STRICT) return fail(STN_ERR_INVALID, "bad mode %d", mode);
    for (int g = 0; g < G; ++g) {
        if (!m->ctx[g]->has_prior) return fail(STN_ERR_INVALID, "stn_multi_ensemble_run needs stn_multi_set_prior first");
        if (!x_dev[g] || !lp_dev[g]) return fail(STN_ERR_INVALID, "null walker copy %d", g);
        for (int b = 0; b < G; ++b)
            if (!m->peer_ok[(size_t)g * G + b])
                return fail(STN_ERR_DEVICE, "devices %d and %d have no peer access", m->ctx[g]->device, m->ctx[b]->device);
    }
    const int64_t H = n_walkers / 2;
    if (plan == 0) plan = pick_plan(m->ctx[0], H);      // from the WHOLE half-ensemble: one summation order whatever the number of devices
    // per device: proposal range of a half-step, scratch, events, acceptance counter
    std::vector<int64_t> k0(G), cnt(G);
    for (int g = 0; g < G; ++g) {
        int64_t lo, hi;
        stn_multi_slice(H, G, g, &lo, &hi);
        k0[g] = lo;
        cnt[g] = hi - lo;
        StnCtx* c = m->ctx[g];
        STN_CUDA(cudaSetDevice(c->device));
        int rc = host_pipe_init(c);
        if (rc) return rc;
        const int64_t need = cnt[g] * P + 2 * cnt[g] + 1;
        if (c->sampler_elems < need) {
            STN_CUDA(cudaDeviceSynchronize());
            if (c->sampler_buf) cudaFree(c->sampler_buf);
            c->sampler_buf = nullptr;
            c->sampler_elems = 0;
            STN_CUDA(cudaMalloc(&c->sampler_buf, (size_t)need * sizeof(double)));
            c->sampler_elems = need;
        }
        for (int i = 0; i < 2; ++i)
            if (!c->sampler_ev[i]) STN_CUDA(cudaEventCreateWithFlags(&c->sampler_ev[i], cudaEventDisableTiming));
        if (!c->sampler_nacc) STN_CUDA(cudaMalloc(&c->sampler_nacc, sizeof(unsigned long long)));
        STN_CUDA(cudaMemsetAsync(c->sampler_nacc, 0, sizeof(unsigned long long), c->streams[0]));
    }
    // One host thread per device issues that device's launches, so the per-half-step driver calls (three
    // launches, G - 1 event waits, one record) of all devices run side by side instead of G-fold in series.
    // The threads meet at a host barrier after every half-step: an event must have been RECORDED by its
    // owner's thread before another thread makes its stream wait for it.
    struct HostBarrier {
        std::atomic<int> count{0};
        std::atomic<int> gen{0};
        int n;
        explicit HostBarrier(int n_) : n(n_) {}
        void wait() {
            const int g = gen.load(std::memory_order_acquire);
            if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
                count.store(0, std::memory_order_relaxed);
                gen.fetch_add(1, std::memory_order_release);
            } else {
                while (gen.load(std::memory_order_acquire) == g) std::this_thread::yield();
            }
        }
    };
    HostBarrier bar(G);
    std::atomic<int> failed{0};
    std::vector<int> rcs(G, STN_OK);
    std::vector<std::string> errs(G);
    auto half_step = [&](int g, int64_t t, int half, int64_t hs) -> int {
        StnCtx* c = m->ctx[g];
        cudaStream_t st = c->streams[0];
        // every device must have finished the previous half-step: its accepted walkers are in this
        // device's copy, and it no longer reads the rows this half-step overwrites
        if (hs > 0)
            for (int b = 0; b < G; ++b)
                if (b != g) STN_CUDA(cudaStreamWaitEvent(st, m->ctx[b]->sampler_ev[(hs - 1) & 1], 0));
        if (cnt[g] > 0) {
            double* y = c->sampler_buf;
            double* z = y + cnt[g] * P;
            double* lpy = z + cnt[g];
            const unsigned grid = (unsigned)((cnt[g] + kThreads - 1) / kThreads);
            (void)cudaGetLastError();
            stn_stretch_propose_part_kernel<<<grid, kThreads, 0, st>>>(x_dev[g], n_walkers, P, ld, half, a, seed,
                                                                       step0 + (uint64_t)t, k0[g], cnt[g], y, z);
            STN_CUDA(cudaGetLastError());
            const int rc = launch_loglike(c, y, cnt[g], P, STN_LAYOUT_AOS, mode, plan, lpy, st, 1, 1);
            if (rc) return rc;
            WalkerCopies C;
            C.n = G;
            C.x[0] = x_dev[g];                   // own copy first: it also supplies the current log-posterior
            C.lp[0] = lp_dev[g];
            for (int b = 0, o = 1; b < G; ++b)
                if (b != g) {
                    C.x[o] = x_dev[b];
                    C.lp[o] = lp_dev[b];
                    ++o;
                }
            WalkerCopies own = C;
            own.n = 1;                           // accept into the local copy ...
            stn_stretch_accept_part_kernel<<<grid, kThreads, 0, st>>>(own, n_walkers, P, ld, half, seed, step0 + (uint64_t)t,
                                                                      k0[g], cnt[g], y, z, lpy, c->sampler_nacc);
            STN_CUDA(cudaGetLastError());
            c->launches += 2;
            if (G > 1) {                         // ... then the slice goes to all other copies, coalesced
                const int64_t row0 = (half ? n_walkers / 2 : 0) + k0[g];
                const int64_t work = cnt[g] * ld / 2 + 1;
                const unsigned bg = (unsigned)((work + kThreads - 1) / kThreads < 4096 ? (work + kThreads - 1) / kThreads : 4096);
                stn_broadcast_rows_kernel<<<bg, kThreads, 0, st>>>(C, row0, cnt[g], ld);
                STN_CUDA(cudaGetLastError());
                c->launches++;
            }
        }
        STN_CUDA(cudaEventRecord(c->sampler_ev[hs & 1], st));
        return STN_OK;
    };
    auto store_chain = [&](int64_t t, int64_t hs) -> int {
        // device 0's copy is complete once every device's last half-step has landed
        StnCtx* c = m->ctx[0];
        cudaStream_t st = c->streams[0];
        for (int b = 1; b < G; ++b) STN_CUDA(cudaStreamWaitEvent(st, m->ctx[b]->sampler_ev[(hs - 1) & 1], 0));
        const int64_t slot = (t + 1) / thin - 1;
        STN_CUDA(cudaMemcpy2DAsync(chain_dev + slot * n_walkers * P, (size_t)P * sizeof(double), x_dev[0],
                                   (size_t)ld * sizeof(double), (size_t)P * sizeof(double), (size_t)n_walkers,
                                   cudaMemcpyDeviceToDevice, st));
        if (lnprob_dev)
            STN_CUDA(cudaMemcpyAsync(lnprob_dev + slot * n_walkers, lp_dev[0], (size_t)n_walkers * sizeof(double),
                                     cudaMemcpyDeviceToDevice, st));
        // the next half-step of the OTHER devices overwrites rows of device 0's copy: they wait for this copy
        STN_CUDA(cudaEventRecord(c->sampler_ev[(hs - 1) & 1], st));
        return STN_OK;
    };
    auto worker = [&](int g) {
        if (cudaSetDevice(m->ctx[g]->device) != cudaSuccess) {
            rcs[g] = STN_ERR_CUDA;
            errs[g] = "cudaSetDevice failed";
            failed.store(1);
        }
        int64_t hs = 0;
        for (int64_t t = 0; t < n_steps; ++t) {
            for (int half = 0; half < 2; ++half, ++hs) {
                if (!failed.load()) {
                    const int rc = half_step(g, t, half, hs);
                    if (rc) {
                        rcs[g] = rc;
                        errs[g] = g_err;
                        failed.store(1);
                    }
                }
                bar.wait();                     // all events of this half-step are recorded
            }
            if (chain_dev && (t + 1) % thin == 0) {
                if (g == 0 && !failed.load()) {
                    const int rc = store_chain(t, hs);
                    if (rc) {
                        rcs[g] = rc;
                        errs[g] = g_err;
                        failed.store(1);
                    }
                }
                bar.wait();                     // device 0's event now covers the copy
            }
        }
    };
    {
        std::vector<std::thread> pool;
        pool.reserve(G - 1);
        for (int g = 1; g < G; ++g) pool.emplace_back(worker, g);
        worker(0);
        for (std::thread& th : pool) th.join();
    }
    int rc = STN_OK;
    for (int g = 0; g < G && !rc; ++g)
        if (rcs[g]) rc = fail(rcs[g], "device %d: %s", m->ctx[g]->device, errs[g].c_str());
    // synchronous on return (also after an error); acceptances summed over the devices
    unsigned long long total = 0;
    for (int g = 0; g < G; ++g) {
        StnCtx* c = m->ctx[g];
        cudaSetDevice(c->device);
        if (c->streams[0]) {
            const cudaError_t e = cudaStreamSynchronize(c->streams[0]);
            if (e != cudaSuccess && !rc) rc = fail(STN_ERR_CUDA, "device %d: %s", c->device, cudaGetErrorString(e));
        }
        unsigned long long v = 0;
        if (!rc && c->sampler_nacc && cudaMemcpy(&v, c->sampler_nacc, sizeof(v), cudaMemcpyDeviceToHost) == cudaSuccess) total += v;
    }
    cudaSetDevice(m->ctx[0]->device);
    if (!rc && n_accept_host) *n_accept_host += (int64_t)total;
    return rc;
}

}  // extern "C"
