// Host side of the marching day kernel as it stood in nesosim_abi.cu (plan builder, eligibility, kernel selection).
// Not compiled: kept with the kernel for the record (see README.md).

typedef void (*MarchKernel)(const DayArgs, const MarchArgs, const StripLink);
MarchKernel march_kernel_of(bool strip, int cpl, int minb) {
    if (cpl == 2) return strip ? day_march_kernel<true, 2, 3> : day_march_kernel<false, 2, 3>;
    if (minb == 5) return strip ? day_march_kernel<true, 1, 5> : day_march_kernel<false, 1, 5>;
    return strip ? day_march_kernel<true, 1, 4> : day_march_kernel<false, 1, 4>;
}
size_t march_smem_bytes(int cpl) { return (cpl == 2 ? sizeof(MarchWarpSmem<2>) : sizeof(MarchWarpSmem<1>)) * MARCH_WARPS; }

void march_release(nesosim_ctx *ctx) {
    cudaFree(ctx->march.rowdesc_dev);
    cudaFree(ctx->march.chains_dev);
    cudaFree(ctx->march.worker_first_dev);
    ctx->march.rowdesc_dev = nullptr;
    ctx->march.chains_dev = nullptr;
    ctx->march.worker_first_dev = nullptr;
    ctx->march.ready = false;
}

// Can the marching kernel run this context's steps?  It carries only what the large-grid configurations use
// (dynamics on, variable density, one shared forcing) and the bare constant divisions, whose operand-window proof
// (cell_math.cuh out_of_guard) needs the same divisor / weight ranges as the season-resident kernel.
bool march_eligible(const nesosim_ctx *ctx) {
    const nesosim_config &c = ctx->cfg;
    if (c.dynamicsInc != 1 || c.density_clim || ctx->member_set_dev) return false;
    if (c.ny < 2 * MARCH_CAP + 1 || c.nx < 4) return false;
    if (!(ctx->g.dx.fast && ctx->g.two_dx.fast && ctx->conv_div.fast)) return false;
    if (!(c.dx >= 1.0 && c.dx <= 536870912.0)) return false;
    if (!(c.deltaT >= 9.5367431640625e-07 && c.deltaT <= 1048576.0)) return false;
    for (int i = 0; i < 9; ++i) {
        const double w = std::fabs(c.conv_weights[i]);
        if (!(w == 0.0 || (w >= 9.5367431640625e-07 && w <= 1048576.0))) return false;
    }
    return std::fabs(c.conv_divisor) >= 9.5367431640625e-07 && std::fabs(c.conv_divisor) <= 1048576.0;
}

// Compile the mask into the marching kernel's plan: per strip and row, what the row needs (closed-form land row, raw
// dynamics, depth / drift rows), and the chains of rows every warp of the launch walks -- cut so that all warps get the
// same modelled cost.  The first and last MARCH_CAP rows of every strip are chains of their own: they hold the grid-edge
// rows (general code) and, for a strip of a decomposed grid, everything that touches the mailboxes; they are dealt
// first, one per warp, so the rows a neighbouring strip waits for leave at the start of the launch.
int march_build(nesosim_ctx *ctx) {
    auto &mp = ctx->march;
    const int has_up = ctx->strip.on ? ctx->strip.has_up : 0, has_dn = ctx->strip.on ? ctx->strip.has_dn : 0;
    if (mp.ready && mp.has_up == has_up && mp.has_dn == has_dn) return NESOSIM_OK;
    march_release(ctx);
    if (const char *env = getenv("NESOSIM_MARCH_CPL")) mp.cpl = atoi(env) == 2 ? 2 : 1;
    mp.minb = mp.cpl == 2 ? 3 : 4;
    if (const char *env = getenv("NESOSIM_MARCH_BLOCKS")) { if (mp.cpl == 1) mp.minb = atoi(env) == 5 ? 5 : 4; }
    const int MW = march_owned(mp.cpl), MRAW = 32 * mp.cpl;
    const int ny = ctx->cfg.ny, nx = ctx->cfg.nx;
    const std::vector<uint8_t> &mask = ctx->mask_host;
    const int ns = (nx + MW - 1) / MW;
    const int ny_pad = ny + 2 * MARCH_PAD;
    std::vector<uint4> rf((size_t)ns * ny_pad, make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u));
    std::vector<double> cost((size_t)ns * ny, 0.0);
    double w_ocean = 300.0, w_land = 128.0, w_row = 600.0;     // modelled cost of an ocean cell, a land cell, a row
    if (const char *env = getenv("NESOSIM_MARCH_COST")) sscanf(env, "%lf,%lf,%lf", &w_ocean, &w_land, &w_row);
    for (int s = 0; s < ns; ++s) {
        const int c0 = s * MW, c1 = std::min(c0 + MW, nx);
        std::vector<uint8_t> oo(ny, 0), raw(ny, 0);
        for (int y = 0; y < ny; ++y) {
            int ocean = 0;
            for (int c = c0; c < c1; ++c) {
                const uint8_t m = mask[(size_t)y * nx + c];
                ocean += !(m > 10 || m < 1);
            }
            oo[y] = ocean > 0;
            cost[(size_t)s * ny + y] = w_row + w_ocean * ocean + w_land * (c1 - c0 - ocean);
        }
        for (int y = 0; y < ny; ++y) raw[y] = oo[y] | (y > 0 ? oo[y - 1] : 0) | (y + 1 < ny ? oo[y + 1] : 0);
        for (int y = 0; y < ny; ++y) {
            const uint8_t h = raw[y] | (y > 0 ? raw[y - 1] : 0) | (y + 1 < ny ? raw[y + 1] : 0);
            unsigned long long bits = 0;                   // land bits of the 64 raw columns c0-1 .. c0+62
            for (int i = 0; i < MRAW; ++i) {
                const int gx = c0 - 1 + i;
                bool land = true;
                if (gx >= 0 && gx < nx) {
                    const uint8_t m = mask[(size_t)y * nx + gx];
                    land = (m > 10 || m < 1);
                }
                if (land) bits |= 1ull << i;
            }
            rf[(size_t)s * ny_pad + MARCH_PAD + y] = make_uint4((unsigned)(bits & 0xffffffffu), (unsigned)(bits >> 32),
                                                                  (unsigned)((oo[y] ? RF_OCEAN : 0) | (raw[y] ? RF_RAW : 0) | (h ? RF_H : 0) | RF_IN), 0u);
        }
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->cfg.device);
    int nctas = sms * mp.minb;
    if (const char *env = getenv("NESOSIM_MARCH_CTAS")) nctas = std::max(1, atoi(env));
    // no more workers than chains of at least ~8 rows would fill
    const long long mid_rows = (long long)ns * std::max(0, ny - 2 * MARCH_CAP);
    nctas = (int)std::max<long long>(1, std::min<long long>(nctas, (mid_rows / 8 + 2 * ns + MARCH_WARPS - 1) / MARCH_WARPS));
    const int nw = nctas * MARCH_WARPS;
    std::vector<std::vector<MarchChain>> per(nw);
    std::vector<double> load(nw, 0.0);
    auto edge_flag = [&](int s, int y0, int y1) {
        const int c0 = s * MW;
        return (c0 - 2 < 0 || c0 + MW + 1 > nx - 1 || y0 - 2 < 0 || y1 + 1 > ny - 1) ? MC_EDGE : 0;
    };
    auto rows_cost = [&](int s, int y0, int y1) {
        double c = 0;
        for (int y = y0; y < y1; ++y) c += cost[(size_t)s * ny + y];
        return c + 2.0 * w_row + 2.0 * w_ocean * MW;      // prologue
    };

    // Cap chains first, one per worker (round-robin); then the rows in between, every worker filled up to the same
    // total.  Every chain costs a prologue, so the target per worker is not known before the cut: start from the ideal
    // and raise it until the greedy cut fits the workers (the last worker, which takes whatever is left, is not overfull).
    const double prologue = 2.0 * w_row + 2.0 * w_ocean * MW;
    double rows_total = 0.0;
    for (size_t i = 0; i < cost.size(); ++i) rows_total += cost[i];
    double target = (rows_total + (2.0 * ns + nw) * prologue) / nw;
    // Grouped cut (default; NESOSIM_MARCH_GROUP=0 gives every warp its own cut): the MARCH_WARPS warps of a CTA take the SAME
    // rows of MARCH_WARPS ADJACENT strips and so walk down side by side -- every plane row is then touched as one
    // contiguous run of MARCH_WARPS strips (2 KB with 62-column strips) within a short time instead of one strip's
    // worth per visit, which is what the DRAM pages and the L2 sector merging of the stores want.
    bool grouped = true;
    if (const char *env = getenv("NESOSIM_MARCH_GROUP")) grouped = atoi(env) != 0;
    if (grouped) {
        const int ng = (ns + MARCH_WARPS - 1) / MARCH_WARPS;
        std::vector<double> gcost((size_t)ng * ny, 0.0);       // a group's row costs what its most expensive strip costs
        for (int s = 0; s < ns; ++s)
            for (int y = 0; y < ny; ++y) gcost[(size_t)(s / MARCH_WARPS) * ny + y] = std::max(gcost[(size_t)(s / MARCH_WARPS) * ny + y], cost[(size_t)s * ny + y]);
        double gtotal = 0.0;
        for (size_t i = 0; i < gcost.size(); ++i) gtotal += gcost[i];
        struct Task { int g, y0, y1; };
        std::vector<std::vector<Task>> tasks(nctas);
        std::vector<double> cl(nctas, 0.0);
        double tgt = (gtotal + (2.0 * ng + nctas) * prologue) / nctas;
        for (int attempt = 0; attempt < 40; ++attempt) {
            for (int i = 0; i < nctas; ++i) { tasks[i].clear(); cl[i] = 0.0; }
            int ccur = 0;
            for (int side = 0; side < 2; ++side)
                for (int g = 0; g < ng; ++g) {
                    const int y0 = side == 0 ? 0 : ny - MARCH_CAP, y1 = side == 0 ? MARCH_CAP : ny;
                    tasks[ccur % nctas].push_back(Task{g, y0, y1});
                    double c = prologue;
                    for (int y = y0; y < y1; ++y) c += gcost[(size_t)g * ny + y];
                    cl[ccur % nctas] += c;
                    ++ccur;
                }
            int c = 0;
            for (int g = 0; g < ng; ++g) {
                int y = MARCH_CAP;
                const int yend = ny - MARCH_CAP;
                while (y < yend) {
                    while (c < nctas - 1 && cl[c] + prologue + gcost[(size_t)g * ny + y] > tgt && cl[c] > 0.0) ++c;
                    int y1 = y;
                    double acc = prologue;
                    while (y1 < yend && (c == nctas - 1 || cl[c] + acc + gcost[(size_t)g * ny + y1] <= tgt || y1 == y)) acc += gcost[(size_t)g * ny + y1++];
                    if (yend - y1 < 3)
                        for (; y1 < yend; ++y1) acc += gcost[(size_t)g * ny + y1];
                    tasks[c].push_back(Task{g, y, y1});
                    cl[c] += acc;
                    y = y1;
                }
            }
            if (cl[nctas - 1] <= tgt * 1.02) break;
            tgt *= 1.02;
        }
        for (int i = 0; i < nw; ++i) { per[i].clear(); load[i] = 0.0; }
        for (int c = 0; c < nctas; ++c)
            for (const Task &t : tasks[c])
                for (int wq = 0; wq < MARCH_WARPS; ++wq) {
                    const int s = t.g * MARCH_WARPS + wq;
                    if (s >= ns) continue;
                    int fl = edge_flag(s, t.y0, t.y1);
                    if (t.y0 == 0 && has_up) fl |= MC_MAIL_TOP | MC_READS_MAIL;
                    if (t.y1 == ny && has_dn) fl |= MC_MAIL_BOT | MC_READS_MAIL;
                    per[c * MARCH_WARPS + wq].push_back(MarchChain{s, t.y0, t.y1, fl});
                    load[c * MARCH_WARPS + wq] += rows_cost(s, t.y0, t.y1);
                }
        target = tgt;
    }
    for (int attempt = 0; attempt < 40 && !grouped; ++attempt) {
        for (int i = 0; i < nw; ++i) { per[i].clear(); load[i] = 0.0; }
        int wcur = 0;
        for (int side = 0; side < 2; ++side)
            for (int s = 0; s < ns; ++s) {
                const int y0 = side == 0 ? 0 : ny - MARCH_CAP, y1 = side == 0 ? MARCH_CAP : ny;
                int fl = edge_flag(s, y0, y1);
                if (side == 0 && has_up) fl |= MC_MAIL_TOP | MC_READS_MAIL;
                if (side == 1 && has_dn) fl |= MC_MAIL_BOT | MC_READS_MAIL;
                per[wcur % nw].push_back(MarchChain{s, y0, y1, fl});
                load[wcur % nw] += rows_cost(s, y0, y1);
                ++wcur;
            }
        int w = 0;
        for (int s = 0; s < ns; ++s) {
            int y = MARCH_CAP;
            const int yend = ny - MARCH_CAP;
            while (y < yend) {
                while (w < nw - 1 && load[w] + prologue + cost[(size_t)s * ny + y] > target && load[w] > 0.0) ++w;
                int y1 = y;
                double c = prologue;
                while (y1 < yend && (w == nw - 1 || load[w] + c + cost[(size_t)s * ny + y1] <= target || y1 == y)) c += cost[(size_t)s * ny + y1++];
                if (yend - y1 < 3) {                 // do not leave a stub of one or two rows behind
                    for (; y1 < yend; ++y1) c += cost[(size_t)s * ny + y1];
                }
                per[w].push_back(MarchChain{s, y, y1, edge_flag(s, y, y1)});
                load[w] += c;
                y = y1;
            }
        }
        if (load[nw - 1] <= target * 1.02) break;
        target *= 1.02;
    }
    std::vector<MarchChain> chains;
    std::vector<int> first(nw + 1, 0);
    for (int i = 0; i < nw; ++i) {
        first[i] = (int)chains.size();
        chains.insert(chains.end(), per[i].begin(), per[i].end());
    }
    first[nw] = (int)chains.size();
    CU(cudaMalloc(&mp.rowdesc_dev, rf.size() * sizeof(uint4)));
    CU(cudaMalloc(&mp.chains_dev, chains.size() * sizeof(MarchChain)));
    CU(cudaMalloc(&mp.worker_first_dev, first.size() * sizeof(int)));
    CU(cudaMemcpy(mp.rowdesc_dev, rf.data(), rf.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(mp.chains_dev, chains.data(), chains.size() * sizeof(MarchChain), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(mp.worker_first_dev, first.data(), first.size() * sizeof(int), cudaMemcpyHostToDevice));
    mp.nstrips = ns;
    mp.nworkers = nw;
    mp.nctas = nctas;
    mp.nchains = (int)chains.size();
    mp.expect_top = has_up ? ns : 0;
    mp.expect_bot = has_dn ? ns : 0;
    mp.has_up = has_up;
    mp.has_dn = has_dn;
    CU(cudaFuncSetAttribute(march_kernel_of(false, mp.cpl, mp.minb), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)march_smem_bytes(mp.cpl)));
    CU(cudaFuncSetAttribute(march_kernel_of(true, mp.cpl, mp.minb), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)march_smem_bytes(mp.cpl)));
    if (getenv("NESOSIM_MARCH_DEBUG")) {
        double mx = 0, mn = 1e300;
        for (int i = 0; i < nw; ++i) { mx = std::max(mx, load[i]); mn = std::min(mn, load[i]); }
        fprintf(stderr, "[march] %d columns per lane, %d strips, %d CTAs, %d workers, %d chains, cost/worker target %.0f min %.0f max %.0f\n",
                mp.cpl, ns, nctas, nw, mp.nchains, target, mn, mx);
    }
    mp.ready = true;
    return NESOSIM_OK;
}

