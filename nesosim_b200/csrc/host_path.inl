// host_path.inl -- nesosim_run_season_host: the end-to-end call with HOST buffers (included by nesosim_abi.cu).
//
// Forcing goes host->device once; members are processed in batches whose device outputs are double-buffered:
// while batch b computes on the compute stream, batch b-1 drains device->host on the copy stream.  Device
// buffers are cached in the context so repeated calls (bench.py) do not re-allocate.
//
// snowAcc and snowOcean (NESOSIM.py:263-270) depend on the forcing only -- pd*C and pd*(1-C) accumulated, no member
// coefficient, no depth -- so with one shared forcing every member's copy is the same array: only member 0's crosses
// PCIe, and host threads replicate it into the other members' slots while the remaining arrays are still draining
// (2 of the 12 arrays per member: 16 % fewer bytes on the link that bounds this call).
namespace {

long long var_elems_per_member(const nesosim_ctx *ctx, int v) {   // v indexes the 11 arrays of nesosim_outputs
    const long long T = ctx->cfg.num_days;
    return (v == 0 ? 2 : 1) * T * ctx->plane;
}

double *const *host_arrays(const nesosim_outputs *o, double *tmp[11]) {
    tmp[0] = o->snowDepths; tmp[1] = o->density; tmp[2] = o->snowAcc; tmp[3] = o->snowOcean; tmp[4] = o->snowAdv;
    tmp[5] = o->snowDiv; tmp[6] = o->snowLead; tmp[7] = o->snowAtm; tmp[8] = o->snowWindPackLoss;
    tmp[9] = o->snowWindPackGain; tmp[10] = o->snowWindPack;
    return tmp;
}

}  // namespace

extern "C" int nesosim_run_season_host(nesosim_ctx *ctx, const double *precip, const double *conc,
                                       const double *wind, const double *drift, const double *rho_clim,
                                       const nesosim_member_params *params, const double *ic, int ic_per_member,
                                       const nesosim_outputs *out_host, int64_t *h2d_bytes, int64_t *d2h_bytes) {
    if (!ctx || !precip || !conc || !wind || !drift || !params || !out_host) return fail(NESOSIM_ERR_ARG, "NULL argument");
    if (ctx->cfg.density_clim && !rho_clim) return fail(NESOSIM_ERR_ARG, "density_clim=1 needs rho_clim");
    CU(cudaSetDevice(ctx->cfg.device));
    // this call copies every batch's results to the host itself, so the season kernel's operand-range flag must be
    // looked at before the copy: asynchronous mode (nesosim_set_async) is suspended for its duration
    struct AsyncOff {
        nesosim_ctx *c;
        bool was;
        explicit AsyncOff(nesosim_ctx *c_) : c(c_), was(c_->async_mode) { c->async_mode = false; }
        ~AsyncOff() { c->async_mode = was; }
    } async_off(ctx);
    if (async_off.was) {
        int rc0 = resolve_pending(ctx, true, nullptr);
        if (rc0) return rc0;
    }
    const int M = ctx->cfg.n_members;
    const long long T = ctx->cfg.num_days, plane = ctx->plane;
    HostPath *hp = &ctx->hp;
    int64_t up = 0, down = 0;

    double *harr[11];
    host_arrays(out_host, harr);
    long long per_member = 0;   // device elements per member for the wanted outputs
    for (int v = 0; v < 11; ++v)
        if (harr[v]) per_member += var_elems_per_member(ctx, v);
    if (per_member == 0) return fail(NESOSIM_ERR_ARG, "no output requested");

    if (!hp->compute) {
        CU(cudaStreamCreateWithFlags(&hp->compute, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&hp->copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&hp->done[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&hp->drained[i], cudaEventDisableTiming));
        }
    }
    const size_t forcing_elems = (size_t)5 * T * plane + T;
    if (!hp->forcing) CU(cudaMalloc(&hp->forcing, forcing_elems * sizeof(double)));
    double *dP = hp->forcing, *dC = dP + T * plane, *dW = dC + T * plane, *dUV = dW + T * plane,
           *dRho = dUV + 2 * T * plane;
    CU(cudaMemcpyAsync(dP, precip, T * plane * 8, cudaMemcpyHostToDevice, hp->compute));
    CU(cudaMemcpyAsync(dC, conc, T * plane * 8, cudaMemcpyHostToDevice, hp->compute));
    CU(cudaMemcpyAsync(dW, wind, T * plane * 8, cudaMemcpyHostToDevice, hp->compute));
    CU(cudaMemcpyAsync(dUV, drift, 2 * T * plane * 8, cudaMemcpyHostToDevice, hp->compute));
    up += 5 * T * plane * 8;
    if (rho_clim) {
        CU(cudaMemcpyAsync(dRho, rho_clim, T * 8, cudaMemcpyHostToDevice, hp->compute));
        up += T * 8;
    }
    const size_t ic_elems = ic ? (size_t)(ic_per_member ? M : 1) * plane : 0;
    if (ic) {
        if (hp->ic_elems < ic_elems) {
            cudaFree(hp->ic);
            hp->ic = nullptr;
            CU(cudaMalloc(&hp->ic, ic_elems * sizeof(double)));
            hp->ic_elems = ic_elems;
        }
        CU(cudaMemcpyAsync(hp->ic, ic, ic_elems * 8, cudaMemcpyHostToDevice, hp->compute));
        up += ic_elems * 8;
    }
    CU(cudaStreamSynchronize(hp->compute));   // set_forcing reads rho_clim back; sources may be pageable
    int rc = nesosim_set_forcing(ctx, dP, dC, dW, dUV, rho_clim ? dRho : nullptr);
    if (rc) return rc;
    if ((rc = upload_coef(ctx, params, hp->compute))) return rc;
    up += (int64_t)sizeof(MemberCoef) * M;

    // batch size: two device buffers of <= NESOSIM_HOST_BATCH_GB (default 16 GiB) each.  Measured on a B200 box
    // (tools/e2e_variants.py, 128 members x 260 days, 21.6 GB to the host): 2 GiB 452 ms, 8 GiB 462 ms, 16 GiB 427 ms --
    // fewer, longer copies keep the link busier (53.9 GB/s is what a plain pinned copy reaches there).
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    double cap_gb = 16.0;
    if (const char *e = getenv("NESOSIM_HOST_BATCH_GB")) cap_gb = atof(e);
    size_t cap = (size_t)(cap_gb * (1ull << 30));
    if (!hp->outbuf[0] && cap * 2 > free_b / 10 * 8) cap = free_b / 10 * 4;
    int batch = (int)std::max<long long>(1, std::min<long long>(M, (long long)(cap / (per_member * 8))));
    // The buffers are cached by their size in BYTES: the bytes a member needs depend on which outputs were asked for,
    // so a later call with more outputs (same batch count) must regrow them.
    const size_t need_bytes = (size_t)batch * per_member * 8;
    if (hp->outbuf_bytes < need_bytes || !hp->outbuf[0] || !hp->outbuf[1]) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(hp->outbuf[i]);
            hp->outbuf[i] = nullptr;
        }
        hp->outbuf_bytes = 0;
        for (int i = 0; i < 2; ++i) {
            cudaError_t e_ = cudaMalloc(&hp->outbuf[i], need_bytes);
            if (e_ != cudaSuccess) {               // leave no half-allocated pair behind
                for (int j = 0; j < 2; ++j) {
                    cudaFree(hp->outbuf[j]);
                    hp->outbuf[j] = nullptr;
                }
                return cuda_fail(e_, "cudaMalloc(output staging)");
            }
        }
        hp->outbuf_bytes = need_bytes;
    }
    batch = (int)std::max<long long>(1, std::min<long long>(batch, (long long)(hp->outbuf_bytes / ((size_t)per_member * 8))));

    // member-independent arrays (v = 2 snowAcc, v = 3 snowOcean): one device->host copy, replicated on the host
    const bool share = M > 1 && ctx->n_sets == 1 && !getenv("NESOSIM_HOST_NO_SHARE");
    if (share && !hp->shared_ready) CU(cudaEventCreateWithFlags(&hp->shared_ready, cudaEventDisableTiming));

    int nb = 0;
    for (int m0 = 0; m0 < M; m0 += batch, ++nb) {
        const int cnt = std::min(batch, M - m0);
        const int b = nb & 1;
        if (nb >= 2) CU(cudaStreamWaitEvent(hp->compute, hp->drained[b], 0));
        // device views of this batch, variable-major inside the buffer
        nesosim_outputs dev{};
        double *darr[11];
        long long off = 0;
        for (int v = 0; v < 11; ++v) {
            darr[v] = nullptr;
            if (!harr[v]) continue;
            darr[v] = hp->outbuf[b] + off;
            off += var_elems_per_member(ctx, v) * cnt;
        }
        dev.snowDepths = darr[0]; dev.density = darr[1]; dev.snowAcc = darr[2]; dev.snowOcean = darr[3];
        dev.snowAdv = darr[4]; dev.snowDiv = darr[5]; dev.snowLead = darr[6]; dev.snowAtm = darr[7];
        dev.snowWindPackLoss = darr[8]; dev.snowWindPackGain = darr[9]; dev.snowWindPack = darr[10];
        dev.depth_member_stride = 2 * T * plane;
        dev.plane_member_stride = T * plane;
        rc = run_members(ctx, ic ? hp->ic : nullptr, ic_per_member, &dev, m0, cnt, 0, -1, hp->compute);
        if (rc) return rc;
        CU(cudaEventRecord(hp->done[b], hp->compute));
        CU(cudaStreamWaitEvent(hp->copy, hp->done[b], 0));
        if (share && m0 == 0) {       // first on the link, so the host threads can start replicating early
            for (int v = 2; v <= 3; ++v) {
                if (!harr[v]) continue;
                const long long n = var_elems_per_member(ctx, v);
                CU(cudaMemcpyAsync(harr[v], darr[v], (size_t)n * 8, cudaMemcpyDeviceToHost, hp->copy));
                down += n * 8;
            }
            CU(cudaEventRecord(hp->shared_ready, hp->copy));
        }
        for (int v = 0; v < 11; ++v) {
            if (!harr[v] || (share && (v == 2 || v == 3))) continue;
            const long long n = var_elems_per_member(ctx, v);
            const long long hstride = (v == 0) ? out_host->depth_member_stride : out_host->plane_member_stride;
            if (hstride == n) {
                CU(cudaMemcpyAsync(harr[v] + (long long)m0 * hstride, darr[v], (size_t)n * cnt * 8, cudaMemcpyDeviceToHost, hp->copy));
            } else {
                for (int m = 0; m < cnt; ++m)
                    CU(cudaMemcpyAsync(harr[v] + (long long)(m0 + m) * hstride, darr[v] + (long long)m * n, (size_t)n * 8,
                                       cudaMemcpyDeviceToHost, hp->copy));
            }
            down += n * cnt * 8;
        }
        CU(cudaEventRecord(hp->drained[b], hp->copy));
    }
    if (share && (harr[2] || harr[3])) {
        CU(cudaEventSynchronize(hp->shared_ready));
        const long long n = var_elems_per_member(ctx, 2), hstride = out_host->plane_member_stride;
        // replication threads: this process' share of the host cores when every visible GPU runs a rank of its own (one
        // process per GPU), at most 16.  The copy into the other members' slots is on the critical path of the call
        // when it is too slow (same box: 2 threads 604 ms, 4 495 ms, 8 462 ms, 16 458 ms; without sharing 480 ms), and
        // 8 ranks x 8 threads oversubscribed the memory bus of a box.  NESOSIM_HOST_THREADS overrides
        int nthreads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency() / (unsigned)std::max(1, nesosim_device_count())));
        if (const char *e = getenv("NESOSIM_HOST_THREADS")) nthreads = std::max(1, atoi(e));
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t)
            pool.emplace_back([=]() {
                for (int m = 1 + t; m < M; m += nthreads)
                    for (int v = 2; v <= 3; ++v)
                        if (harr[v]) std::memcpy(harr[v] + (long long)m * hstride, harr[v], (size_t)n * 8);
            });
        for (auto &th : pool) th.join();
    }
    CU(cudaStreamSynchronize(hp->compute));
    CU(cudaStreamSynchronize(hp->copy));
    if (h2d_bytes) *h2d_bytes = up;
    if (d2h_bytes) *d2h_bytes = down;
    return NESOSIM_OK;
}
