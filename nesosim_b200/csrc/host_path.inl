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
//
// Compacted drain (drain_kernels.cuh; NESOSIM_HOST_COMPACT=0/1 overrides the choice): of the nine member-dependent arrays
// only the ocean cells (43 % of the polar grid) and the land cells of the first three time slots cross the link, in
// chunks of a few members that land in a ring of pinned slots; a pool of host threads scatters every chunk into the
// caller's arrays -- plane by plane through a cache-resident scratch plane and out with streaming stores -- while the
// next chunks are on the link.  A chunk whose land cells turn out not to be constant is copied in full instead.
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
namespace {

// dst <- src (n doubles) with non-temporal stores: the caller's arrays are written once and not read by this call, so
// going around the cache saves the read-for-ownership of every line (half of the memory traffic of the scatter)
void stream_plane(double *dst, const double *src, long long n) {
#if defined(__SSE2__)
    long long i = 0;
    if (((uintptr_t)dst & 15) && n) { dst[0] = src[0]; i = 1; }
    for (; i + 8 <= n; i += 8) {
        _mm_stream_pd(dst + i, _mm_loadu_pd(src + i));
        _mm_stream_pd(dst + i + 2, _mm_loadu_pd(src + i + 2));
        _mm_stream_pd(dst + i + 4, _mm_loadu_pd(src + i + 4));
        _mm_stream_pd(dst + i + 6, _mm_loadu_pd(src + i + 6));
    }
    for (; i + 2 <= n; i += 2) _mm_stream_pd(dst + i, _mm_loadu_pd(src + i));
    if (i < n) dst[i] = src[i];
#else
    std::memcpy(dst, src, (size_t)n * 8);
#endif
}

// One member's packed block of one array -> the caller's full planes.  oc: [pps*T][n_ocean]; the land cells of the
// first `head` slots follow; a later slot repeats the land cells of slot head-1 (they stay in the scratch plane).
void scatter_member_array(const std::vector<int> &ocean_idx, const std::vector<int> &land_idx, long long plane, int pps,
                          int T, int head, const double *oc, double *dst, double *scratch) {
    const long long n_ocean = (long long)ocean_idx.size(), n_land = (long long)land_idx.size();
    const double *ld = oc + (long long)pps * T * n_ocean;
    const int *oi = ocean_idx.data(), *li = land_idx.data();
    for (int layer = 0; layer < pps; ++layer)
        for (int slot = 0; slot < T; ++slot) {
            const long long q = (long long)slot * pps + layer;
            if (slot < head) {
                const double *l = ld + q * n_land;
                for (long long i = 0; i < n_land; ++i) scratch[li[i]] = l[i];
            }
            const double *o = oc + q * n_ocean;
            for (long long i = 0; i < n_ocean; ++i) scratch[oi[i]] = o[i];
            stream_plane(dst + q * plane, scratch, plane);
        }
#if defined(__SSE2__)
    _mm_sfence();
#endif
}

// Host threads of one call: a queue of closures; the main thread keeps the link busy and hands out the work.
struct DrainPool {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void(double *)>> q;
    std::vector<std::thread> th;
    bool stop = false;
    void start(int n, size_t scratch_elems) {
        for (int t = 0; t < n; ++t)
            th.emplace_back([this, scratch_elems]() {
                std::vector<double> scratch(scratch_elems);
                for (;;) {
                    std::function<void(double *)> f;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [this] { return stop || !q.empty(); });
                        if (q.empty()) return;
                        f = std::move(q.front());
                        q.pop_front();
                    }
                    f(scratch.data());
                }
            });
    }
    void push(std::function<void(double *)> f) {
        {
            std::lock_guard<std::mutex> lk(mu);
            q.push_back(std::move(f));
        }
        cv.notify_one();
    }
    void finish() {   // run the queue dry, then join
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto &t : th) t.join();
        th.clear();
    }
    ~DrainPool() { finish(); }
};

int host_threads() {
    // this process' share of the host cores when every visible GPU runs a rank of its own (one process per GPU), at most
    // 16; NESOSIM_HOST_THREADS overrides (bench.py sets it to cores / world size)
    int n = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency() / (unsigned)std::max(1, nesosim_device_count())));
    if (const char *e = getenv("NESOSIM_HOST_THREADS")) n = std::max(1, atoi(e));
    return n;
}

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

// The host half of the compacted drain on its own (no device involved): one member's packed block of one array ->
// full planes.  What nesosim_run_season_host's threads run; exported so that the layout has a test without a GPU.
extern "C" int nesosim_unpack_member_array(const uint8_t *mask, int64_t plane, int planes_per_slot, int num_days,
                                           const double *packed, double *dst) {
    if (!mask || !packed || !dst || plane < 1 || planes_per_slot < 1 || num_days < 1) return fail(NESOSIM_ERR_ARG, "bad argument");
    std::vector<int> oi, li;
    for (int64_t c = 0; c < plane; ++c) ((mask[c] > 10 || mask[c] < 1) ? li : oi).push_back((int)c);
    std::vector<double> scratch((size_t)plane);
    scatter_member_array(oi, li, plane, planes_per_slot, num_days, (int)std::min<long long>(DRAIN_HEAD, num_days), packed, dst, scratch.data());
    return NESOSIM_OK;
}

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

    // ---- which drain?  The compacted one pays when the link is slower than this process' host threads can scatter.
    const int nthreads = host_threads();
    const bool share = M > 1 && ctx->n_sets == 1 && !getenv("NESOSIM_HOST_NO_SHARE");
    if (hp->n_ocean < 0) {
        for (long long c = 0; c < plane; ++c) {
            const uint8_t mk = ctx->mask_host[(size_t)c];
            ((mk > 10 || mk < 1) ? hp->land_idx : hp->ocean_idx).push_back((int)c);
        }
        hp->n_ocean = (int)hp->ocean_idx.size();
        hp->n_land = (int)hp->land_idx.size();
    }
    int carr[DRAIN_MAX_ARRAYS], n_carr = 0;        // the member-dependent arrays that were asked for
    for (int v = 0; v < 11; ++v)
        if (harr[v] && v != 2 && v != 3) carr[n_carr++] = v;
    // Measured (128 members x 260 days per GPU, 21.6 GB of arrays; tools/e2e_variants.py, tools/e2e_multi_gpu.py;
    // profiles/r02_e2e_compacted_drain.jsonl).  One rank on a 16-core box: plain drain 425-455 ms (the link: 54 GB/s);
    // every block packed: 1186 ms with 2 host threads, 665 with 4, 372 with 8, 325-336 with 12 or 16 -- then the host's
    // memory system binds (the scatter writes 26 GB, and the ring is written and read once more); hybrid (below): 408-452
    // with 2 threads, 356 with 4, 334-353 with 8, 330 with 16 -- never behind the plain drain.  That memory system is
    // shared by the ranks of a box: two ranks on a 24-core box 418 ms plain, 502 all packed, 420 hybrid with 12 threads
    // each, 402 with 6; eight ranks on a 32-core box 2168 ms plain, 2202 all packed (hybrid not measured there).
    // So: the compacted (hybrid) drain from 6 host threads per rank on -- cores / visible GPUs by default, which keeps
    // the 8-GPU box on the plain drain it was measured with.
    bool compact = n_carr > 0 && T > DRAIN_HEAD && plane < (1ll << 31) && nthreads >= 6 && (double)hp->n_ocean <= 0.6 * (double)plane;
    if (const char *e = getenv("NESOSIM_HOST_COMPACT")) compact = n_carr > 0 && plane < (1ll << 31) && atoi(e) != 0;
    int head = (int)std::min<long long>(DRAIN_HEAD, T);
    // (tests only: fewer head slots than the model needs, so that the land-cell check trips and the chunk is copied in full)
    if (const char *e = getenv("NESOSIM_DRAIN_HEAD")) head = std::max(1, std::min(head, atoi(e)));
    long long rec_elems = 0, rec_off[DRAIN_MAX_ARRAYS];
    for (int i = 0; i < n_carr; ++i) {
        const int pps = carr[i] == 0 ? 2 : 1;
        rec_off[i] = rec_elems;
        rec_elems += (long long)pps * T * hp->n_ocean + (long long)pps * head * hp->n_land;
    }

    // batch size: two device buffers of <= NESOSIM_HOST_BATCH_GB each (default 16 GiB; 4 GiB for the compacted drain,
    // whose first chunk cannot leave before the first batch is computed and packed).  Measured on a B200 box
    // (tools/e2e_variants.py, 128 members x 260 days, 21.6 GB to the host, full drain): 2 GiB 452 ms, 8 GiB 462 ms,
    // 16 GiB 427 ms -- fewer, longer copies keep the link busier (53.9 GB/s is what a plain pinned copy reaches there).
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    double cap_gb = compact ? 4.0 : 16.0;
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
    const int n_batches = (M + batch - 1) / batch;

    // compacted drain: the packed blocks -- one per (member, array), contiguous in member-major order -- travel in chunks
    // of NESOSIM_HOST_CHUNK_MB (default 32) into a ring of NESOSIM_HOST_RING (default 6) pinned slots
    int RING = 6;
    if (const char *e = getenv("NESOSIM_HOST_RING")) RING = std::max(2, std::min(32, atoi(e)));
    long long chunk_target = 32ll << 20;
    if (const char *e = getenv("NESOSIM_HOST_CHUNK_MB")) chunk_target = (long long)(atof(e) * (1 << 20));
    struct Chunk {
        int nb;                 // batch
        int t0, t1;             // blocks [t0, t1) of the batch; block t = member-in-batch * n_carr + array
        long long off, elems;   // inside the batch's packed buffer
    };
    std::vector<Chunk> chunks;             // formed as the drain proceeds (see the pipeline below)
    long long max_chunk_elems = 0;
    int max_chunk_blocks = 0, total_blocks = 0;
    auto block_elems = [&](int i) { return (i + 1 < n_carr ? rec_off[i + 1] : rec_elems) - rec_off[i]; };
    // hybrid drain: while the host threads are the bottleneck (every ring slot taken), the copy engine moves whole
    // (member, array) blocks of the same batch straight into the caller's arrays, from the far end of the batch
    bool hybrid = compact;
    if (const char *e = getenv("NESOSIM_HOST_HYBRID")) hybrid = compact && atoi(e) != 0;
    if (compact) {
        long long min_blk = block_elems(0), max_blk = block_elems(0);
        for (int i = 1; i < n_carr; ++i) {
            min_blk = std::min(min_blk, block_elems(i));
            max_blk = std::max(max_blk, block_elems(i));
        }
        max_chunk_elems = std::max(max_blk, chunk_target / 8);
        max_chunk_blocks = (int)std::min<long long>((long long)batch * n_carr, chunk_target / 8 / std::max<long long>(1, min_blk) + 1);
        total_blocks = M * n_carr;
        chunks.reserve((size_t)total_blocks);
        if (!hp->cells_dev) {
            CU(cudaMalloc(&hp->cells_dev, (size_t)plane * sizeof(int)));
            CU(cudaMemcpy(hp->cells_dev, hp->ocean_idx.data(), (size_t)hp->n_ocean * sizeof(int), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(hp->cells_dev + hp->n_ocean, hp->land_idx.data(), (size_t)hp->n_land * sizeof(int), cudaMemcpyHostToDevice));
        }
        const size_t pk = (size_t)batch * rec_elems * 8;
        if (hp->packed_bytes < pk) {
            for (int i = 0; i < 2; ++i) {
                cudaFree(hp->packed[i]);
                hp->packed[i] = nullptr;
            }
            hp->packed_bytes = 0;
            for (int i = 0; i < 2; ++i) CU(cudaMalloc(&hp->packed[i], pk));
            hp->packed_bytes = pk;
        }
        // a ring slot: the chunk's blocks, then one counter per block
        const size_t slot_bytes = (size_t)max_chunk_elems * 8 + (size_t)max_chunk_blocks * 8;
        if (hp->ring_bytes < RING * slot_bytes) {
            if (hp->ring) cudaFreeHost(hp->ring);
            hp->ring = nullptr;
            hp->ring_bytes = 0;
            CU(cudaHostAlloc((void **)&hp->ring, RING * slot_bytes, cudaHostAllocDefault));
            hp->ring_bytes = RING * slot_bytes;
        }
        while ((int)hp->arrived.size() < RING) {
            cudaEvent_t ev;
            CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            hp->arrived.push_back(ev);
        }
        if (hp->chunk_flags_n < (size_t)2 * batch * n_carr) {
            cudaFree(hp->chunk_flags);
            hp->chunk_flags = nullptr;
            hp->chunk_flags_n = 0;
            CU(cudaMalloc(&hp->chunk_flags, (size_t)2 * batch * n_carr * sizeof(unsigned long long)));
            hp->chunk_flags_n = (size_t)2 * batch * n_carr;
        }
    }
    const size_t slot_bytes = (size_t)max_chunk_elems * 8 + (size_t)max_chunk_blocks * 8;
    if (share && !hp->shared_ready) CU(cudaEventCreateWithFlags(&hp->shared_ready, cudaEventDisableTiming));

    // progress of the scatter, one counter per chunk of the call (guarded by prog_mu); declared before the pool, whose
    // destructor runs the queue dry
    std::mutex prog_mu;
    std::vector<int> pending((size_t)total_blocks + 1, 0);   // (a chunk holds at least one block)
    DrainPool pool;
    const bool need_pool = compact || (share && (harr[2] || harr[3]));
    if (need_pool) pool.start(nthreads, compact ? (size_t)plane : 1);

    struct BatchView {
        int m0 = 0, cnt = 0;
        double *darr[11];
    };
    std::vector<BatchView> views(n_batches);

    // ---- compute (and pack) of batch nb, on the compute stream
    auto launch = [&](int nb) -> int {
        BatchView &bv = views[nb];
        bv.m0 = nb * batch;
        bv.cnt = std::min(batch, M - bv.m0);
        const int b = nb & 1;
        if (nb >= 2 && !compact) CU(cudaStreamWaitEvent(hp->compute, hp->drained[b], 0));
        // device views of this batch, variable-major inside the buffer
        nesosim_outputs dev{};
        long long off = 0;
        for (int v = 0; v < 11; ++v) {
            bv.darr[v] = nullptr;
            if (!harr[v]) continue;
            bv.darr[v] = hp->outbuf[b] + off;
            off += var_elems_per_member(ctx, v) * bv.cnt;
        }
        double **darr = bv.darr;
        dev.snowDepths = darr[0]; dev.density = darr[1]; dev.snowAcc = darr[2]; dev.snowOcean = darr[3];
        dev.snowAdv = darr[4]; dev.snowDiv = darr[5]; dev.snowLead = darr[6]; dev.snowAtm = darr[7];
        dev.snowWindPackLoss = darr[8]; dev.snowWindPackGain = darr[9]; dev.snowWindPack = darr[10];
        dev.depth_member_stride = 2 * T * plane;
        dev.plane_member_stride = T * plane;
        int rc_ = run_members(ctx, ic ? hp->ic : nullptr, ic_per_member, &dev, bv.m0, bv.cnt, 0, -1, hp->compute);
        if (rc_) return rc_;
        if (compact) {
            PackArgs pa{};
            pa.n_arrays = n_carr; pa.T = (int)T; pa.head = head;
            pa.plane = (int)plane; pa.n_ocean = hp->n_ocean; pa.n_land = hp->n_land;
            pa.ocean_idx = hp->cells_dev; pa.land_idx = hp->cells_dev + hp->n_ocean;
            pa.rec_elems = rec_elems;
            pa.plane0[0] = 0;
            for (int i = 0; i < n_carr; ++i) {
                const int v = carr[i];
                pa.pps[i] = v == 0 ? 2 : 1;
                pa.mstride[i] = var_elems_per_member(ctx, v);
                pa.rec_off[i] = rec_off[i];
                pa.plane0[i + 1] = pa.plane0[i] + pa.pps[i] * (int)T;
                pa.src[i] = darr[v];
            }
            pa.dst = hp->packed[b];
            pa.flag = hp->chunk_flags + (size_t)b * batch * n_carr;
            CU(cudaMemsetAsync(pa.flag, 0, (size_t)bv.cnt * n_carr * sizeof(unsigned long long), hp->compute));
            pack_ocean_kernel<<<dim3((unsigned)pa.plane0[n_carr], (unsigned)bv.cnt), 256, 0, hp->compute>>>(pa);
            CU(cudaGetLastError());
            ctx->launches += 1;
        }
        CU(cudaEventRecord(hp->done[b], hp->compute));
        return NESOSIM_OK;
    };

    // ---- member-independent arrays (v = 2 snowAcc, v = 3 snowOcean): one device->host copy, replicated on the host
    auto drain_shared = [&](const BatchView &bv) -> int {
        for (int v = 2; v <= 3; ++v) {
            if (!harr[v]) continue;
            const long long n = var_elems_per_member(ctx, v);
            CU(cudaMemcpyAsync(harr[v], bv.darr[v], (size_t)n * 8, cudaMemcpyDeviceToHost, hp->copy));
            down += n * 8;
        }
        CU(cudaEventRecord(hp->shared_ready, hp->copy));
        return NESOSIM_OK;
    };
    auto copy_full = [&](const BatchView &bv, int v, int ml, int cnt) -> int {   // members ml .. ml+cnt-1 of the batch
        const long long n = var_elems_per_member(ctx, v);
        const long long hstride = (v == 0) ? out_host->depth_member_stride : out_host->plane_member_stride;
        if (hstride == n) {
            CU(cudaMemcpyAsync(harr[v] + (long long)(bv.m0 + ml) * hstride, bv.darr[v] + (long long)ml * n, (size_t)n * cnt * 8,
                               cudaMemcpyDeviceToHost, hp->copy));
        } else {
            for (int m = ml; m < ml + cnt; ++m)
                CU(cudaMemcpyAsync(harr[v] + (long long)(bv.m0 + m) * hstride, bv.darr[v] + (long long)m * n, (size_t)n * 8,
                                   cudaMemcpyDeviceToHost, hp->copy));
        }
        down += n * cnt * 8;
        return NESOSIM_OK;
    };
    bool replication_queued = false;
    auto queue_replication = [&]() {
        if (replication_queued || !share || !(harr[2] || harr[3])) return;
        replication_queued = true;
        const long long n = var_elems_per_member(ctx, 2), hstride = out_host->plane_member_stride;
        for (int m = 1; m < M; ++m)
            for (int v = 2; v <= 3; ++v)
                if (harr[v]) {
                    double *dst = harr[v] + (long long)m * hstride;
                    const double *src = harr[v];
                    pool.push([=](double *) { stream_plane(dst, src, n); });
                }
    };

    if (!compact) {
        // ---- plain drain: batch nb drains on the copy stream while batch nb+1 computes
        for (int nb = 0; nb < n_batches; ++nb) {
            if ((rc = launch(nb))) return rc;
            const BatchView &bv = views[nb];
            const int b = nb & 1;
            CU(cudaStreamWaitEvent(hp->copy, hp->done[b], 0));
            if (share && nb == 0 && (rc = drain_shared(bv))) return rc;   // first on the link: replication starts early
            for (int v = 0; v < 11; ++v) {
                if (!harr[v] || (share && (v == 2 || v == 3))) continue;
                if ((rc = copy_full(bv, v, 0, bv.cnt))) return rc;
            }
            CU(cudaEventRecord(hp->drained[b], hp->copy));
        }
    } else {
        // ---- compacted drain: one pipeline over all the blocks of the call.  The main thread launches batches as their
        // buffers come free, keeps up to RING chunk copies queued on the link, hands every arrived chunk to the pool,
        // and -- hybrid -- gives the link whole blocks to copy whenever the ring is full.
        struct BatchDrain {
            int lo = 0, hi = 0;         // blocks [lo, hi) not yet issued: packed chunks take from lo, plain copies from hi
            int last_chunk = -1;        // index of the batch's last chunk (or of the last chunk before it)
            int plain_out = 0;          // plain copies of the batch that have not completed
            bool issued = false;
        };
        std::vector<BatchDrain> bd(n_batches);
        constexpr int PLAIN_MAX = 2;
        cudaEvent_t plain_ev[PLAIN_MAX];
        int plain_nb[PLAIN_MAX], plain_n = 0, plain_head = 0;      // FIFO of plain copies in flight
        for (int i = 0; i < PLAIN_MAX; ++i) {
            while ((int)hp->arrived.size() < RING + PLAIN_MAX) {
                cudaEvent_t ev;
                CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                hp->arrived.push_back(ev);
            }
            plain_ev[i] = hp->arrived[RING + i];
        }
        int launched = 0, cur = 0, taken = 0;
        long long n_plain = 0, n_packed = 0;
        auto slot_free = [&](int ch) {  // has the chunk that used this ring slot before been scattered?
            if (ch < RING) return true;
            std::lock_guard<std::mutex> lk(prog_mu);
            return pending[ch - RING] == 0;
        };
        auto batch_drained = [&](int nb) {   // every copy out of the batch's device buffers has completed
            return bd[nb].issued && taken > bd[nb].last_chunk && bd[nb].plain_out == 0;
        };
        while (cur < n_batches || taken < (int)chunks.size() || plain_n > 0) {
            bool progress = false;
            if (launched < n_batches && (launched < 2 || batch_drained(launched - 2))) {
                if ((rc = launch(launched))) return rc;
                const BatchView &bv = views[launched];
                bd[launched].hi = bv.cnt * n_carr;
                CU(cudaStreamWaitEvent(hp->copy, hp->done[launched & 1], 0));
                if (launched == 0 && share && (harr[2] || harr[3]) && (rc = drain_shared(bv))) return rc;
                if (!share)
                    for (int v = 2; v <= 3; ++v)
                        if (harr[v]) {      // not shared: every member's copy crosses the link; tracked like a plain block copy
                            if ((rc = copy_full(bv, v, 0, bv.cnt))) return rc;
                        }
                ++launched;
                progress = true;
            }
            if (cur < launched) {
                BatchDrain &d = bd[cur];
                const BatchView &bv = views[cur];
                const int b = cur & 1, issued = (int)chunks.size();
                if (d.lo == d.hi) {
                    d.issued = true;
                    d.last_chunk = issued - 1;
                    ++cur;
                    progress = true;
                } else if (issued < taken + RING && slot_free(issued)) {
                    Chunk c{cur, d.lo, d.lo, (long long)(d.lo / n_carr) * rec_elems + rec_off[d.lo % n_carr], 0};
                    do {
                        c.elems += block_elems(c.t1 % n_carr);
                        ++c.t1;
                    } while (c.t1 < d.hi && (c.elems + block_elems(c.t1 % n_carr)) * 8 <= chunk_target && c.t1 - c.t0 < max_chunk_blocks);
                    d.lo = c.t1;
                    char *slot = (char *)hp->ring + (size_t)(issued % RING) * slot_bytes;
                    CU(cudaMemcpyAsync(slot, hp->packed[b] + c.off, (size_t)c.elems * 8, cudaMemcpyDeviceToHost, hp->copy));
                    CU(cudaMemcpyAsync(slot + (size_t)max_chunk_elems * 8, hp->chunk_flags + (size_t)b * batch * n_carr + c.t0,
                                       (size_t)(c.t1 - c.t0) * 8, cudaMemcpyDeviceToHost, hp->copy));
                    CU(cudaEventRecord(hp->arrived[issued % RING], hp->copy));
                    down += (int64_t)c.elems * 8 + (c.t1 - c.t0) * 8;
                    n_packed += c.t1 - c.t0;
                    chunks.push_back(c);
                    progress = true;
                } else if (hybrid && plain_n < PLAIN_MAX) {
                    const int t = --d.hi, slot_i = (plain_head + plain_n) % PLAIN_MAX;
                    if ((rc = copy_full(bv, carr[t % n_carr], t / n_carr, 1))) return rc;
                    CU(cudaEventRecord(plain_ev[slot_i], hp->copy));
                    plain_nb[slot_i] = cur;
                    ++plain_n;
                    ++d.plain_out;
                    ++n_plain;
                    progress = true;
                }
            }
            if (plain_n > 0) {
                const cudaError_t q_ = cudaEventQuery(plain_ev[plain_head]);
                if (q_ == cudaSuccess) {
                    --bd[plain_nb[plain_head]].plain_out;
                    plain_head = (plain_head + 1) % PLAIN_MAX;
                    --plain_n;
                    progress = true;
                } else if (q_ != cudaErrorNotReady) {
                    return cuda_fail(q_, "device->host copy of a block");
                }
            }
            if (share && (harr[2] || harr[3]) && !replication_queued && cudaEventQuery(hp->shared_ready) == cudaSuccess) {
                queue_replication();
                progress = true;
            }
            if (taken < (int)chunks.size()) {
                const cudaError_t q_ = cudaEventQuery(hp->arrived[taken % RING]);
                if (q_ == cudaSuccess) {
                    const Chunk c = chunks[taken];
                    const BatchView &bv = views[c.nb];
                    const char *slot = (const char *)hp->ring + (size_t)(taken % RING) * slot_bytes;
                    const unsigned long long *bad = (const unsigned long long *)(slot + (size_t)max_chunk_elems * 8);
                    int n_tasks = 0;
                    for (int t = c.t0; t < c.t1; ++t) n_tasks += bad[t - c.t0] == 0;
                    {
                        std::lock_guard<std::mutex> lk(prog_mu);
                        pending[taken] = n_tasks;
                    }
                    long long boff = 0;      // of the block inside the slot
                    for (int t = c.t0; t < c.t1; ++t) {
                        const int ml = t / n_carr, i = t % n_carr, v = carr[i], pps = v == 0 ? 2 : 1;
                        const double *oc = (const double *)slot + boff;
                        boff += block_elems(i);
                        if (bad[t - c.t0]) {     // land cells not constant: the plain copy of this member's array
                            if ((rc = copy_full(bv, v, ml, 1))) return rc;
                            // (rare path: waited for on the spot, the batch's buffers are reused two batches on)
                            CU(cudaStreamSynchronize(hp->copy));
                            ctx->hp_full_chunks++;
                            continue;
                        }
                        const long long hstride = (v == 0) ? out_host->depth_member_stride : out_host->plane_member_stride;
                        double *dst = harr[v] + (long long)(bv.m0 + ml) * hstride;
                        int *cnt_p = &pending[taken];
                        std::mutex *mu_p = &prog_mu;
                        const std::vector<int> *oi = &hp->ocean_idx, *li = &hp->land_idx;
                        const int T_ = (int)T;
                        pool.push([=](double *scratch) {
                            scatter_member_array(*oi, *li, plane, pps, T_, head, oc, dst, scratch);
                            std::lock_guard<std::mutex> lk(*mu_p);
                            --*cnt_p;
                        });
                    }
                    ++taken;
                    progress = true;
                } else if (q_ != cudaErrorNotReady) {
                    return cuda_fail(q_, "device->host copy of a packed chunk");
                }
            }
            if (!progress) std::this_thread::sleep_for(std::chrono::microseconds(10));
        }
        ctx->hp_blocks_packed = n_packed;
        ctx->hp_blocks_plain = n_plain;
    }
    if (share && (harr[2] || harr[3])) {
        CU(cudaEventSynchronize(hp->shared_ready));
        queue_replication();
    }
    pool.finish();
    CU(cudaStreamSynchronize(hp->compute));
    CU(cudaStreamSynchronize(hp->copy));
    if (h2d_bytes) *h2d_bytes = up;
    if (d2h_bytes) *d2h_bytes = down;
    ctx->hp_last_compact = compact ? 1 : 0;
    return NESOSIM_OK;
}
