/* nesosim_b200.h -- C ABI of the B200-native NESOSIM daily snow-budget path.
 *
 * Drop-in boundary for ONE hot path of ac137/NESOSIM: `calcBudget` and its callees
 * (reference: source/NESOSIM.py:51-347,458-473) plus the day loop of `main` that drives it
 * (source/NESOSIM.py:586-649).  The reference is pure Python, so there is no existing FFI; these are the
 * entry points a ctypes binding for that path binds (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *   - plain C, no torch/CUDA types in signatures; `stream` is a cudaStream_t passed as void* (NULL = default).
 *   - every `*_dev` / "device pointer" argument is a CUDA device address (e.g. torch.Tensor.data_ptr());
 *     arguments documented "host" are ordinary host memory.
 *   - all fields are float64 unless stated; planes are dense row-major (ny rows, nx columns), time-major
 *     stacks are [T][ny][nx] exactly like the reference's numpy arrays (genEmptyArrays, NESOSIM.py:350-376).
 *   - functions return 0 (NESOSIM_OK) or a negative error code; nesosim_last_error() gives the message for
 *     the calling thread.  NaN/inf in the data are data, not errors (NESOSIM.py relies on NaN algebra).
 *   - the caller owns every buffer; the library allocates only an opaque context (and optional scratch).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns NESOSIM_ERR_CUDA.
 */
#ifndef NESOSIM_B200_H
#define NESOSIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NESOSIM_ABI_VERSION 1

#define NESOSIM_OK            0
#define NESOSIM_ERR_ARG      -1   /* bad argument (NULL, non-positive size, ny or nx < 2, ...) */
#define NESOSIM_ERR_CUDA     -2   /* CUDA runtime error or no device */
#define NESOSIM_ERR_STATE    -3   /* call order (e.g. run_season before set_forcing) */
#define NESOSIM_ERR_NOMEM    -4

typedef struct nesosim_ctx nesosim_ctx;

/* Model constants `main` assigns to module globals (NESOSIM.py:527-541) + grid + switches of calcBudget
 * (NESOSIM.py:227).  conv_weights is `Gaussian2DKernel(x_stddev=1,x_size=3,y_size=3).array` (row-major 3x3)
 * and conv_divisor its `.sum()`, both computed by the HOST with numpy so the device uses bit-identical
 * constants (smooth_snow, NESOSIM.py:170-187); pass weights/sum and divisor 1.0 for a pre-normalised kernel. */
typedef struct nesosim_config {
    int32_t ny, nx;              /* grid rows / columns (>= 2 each: np.gradient needs two points)        */
    int32_t num_days;            /* T = numDays: number of time slots; the season has T-1 steps          */
    int32_t n_members;           /* M parameter sets sharing one forcing (1 for a plain `main` run)      */
    double  dx;                  /* grid spacing [m]: the `dx` argument of main / calcDynamics           */
    double  deltaT;              /* 86400.                                                               */
    double  snowDensityFresh;    /* 200.                                                                 */
    double  snowDensityOld;      /* 350.                                                                 */
    double  minSnowD;            /* 0.02                                                                 */
    double  minConc;             /* 0.15                                                                 */
    double  conv_weights[9];
    double  conv_divisor;
    int32_t dynamicsInc, leadlossInc, windpackInc, atmlossInc;
    int32_t density_clim;        /* 0: densityType='variable', 1: 'clim' (needs rho_clim)                */
    int32_t device;              /* CUDA device ordinal                                                  */
} nesosim_config;

/* Per-member coefficients: windPackFactorT, windPackThreshT, leadLossFactorT, atmLossFactorT of main. */
typedef struct nesosim_member_params {
    double windPackFactor, windPackThresh, leadLossFactor, atmLossFactor;
} nesosim_member_params;

/* The twelve arrays calcBudget advances (NESOSIM.py:264-347).  Device pointers; member m of variable v
 * starts at v + m*member_stride elements; within a member the layout is the reference's:
 * snowDepths [T][2][ny][nx], everything else [T][ny][nx].  NULL = not wanted (kept in internal scratch). */
typedef struct nesosim_outputs {
    double *snowDepths;
    double *density;
    double *snowAcc, *snowOcean, *snowAdv, *snowDiv, *snowLead, *snowAtm;
    double *snowWindPackLoss, *snowWindPackGain, *snowWindPack;
    int64_t depth_member_stride;   /* elements between members in snowDepths (>= T*2*ny*nx)             */
    int64_t plane_member_stride;   /* elements between members in every other array (>= T*ny*nx)         */
} nesosim_outputs;

int         nesosim_abi_version(void);
const char *nesosim_last_error(void);
/* Number of CUDA devices visible (0 on a CPU-only host; never fails). */
int         nesosim_device_count(void);

/* Context = grid, constants, region mask.  `region_mask_host` is ny*nx uint8 region codes (host memory);
 * only `>10` (land/coast) and `<1` (lakes) are ever tested (fill_nan_no_negative, NESOSIM.py:158-162). */
int nesosim_create(const nesosim_config *cfg, const uint8_t *region_mask_host, nesosim_ctx **out);
int nesosim_destroy(nesosim_ctx *ctx);

/* Register a season of forcing already resident in HBM (what loadData returns per day, NESOSIM.py:379-456,
 * stacked over T days): precip/conc/wind [T][ny][nx], drift [T][2][ny][nx], rho_clim [T] (fresh-snow density
 * per step for density_clim=1, else NULL).  Pointers are borrowed, not copied. */
int nesosim_set_forcing(nesosim_ctx *ctx, const double *precip_dev, const double *conc_dev,
                        const double *wind_dev, const double *drift_dev, const double *rho_clim_dev);

/* A batch of seasons in one context (the loop of run_multiseason.py:39-50 as one native call): `n_sets` independent
 * forcing stacks laid out [n_sets][T][ny][nx] (drift [n_sets][T][2][ny][nx]); member m runs on stack
 * member_set_host[m] for set_days_host[set] days (2 <= days <= T: seasons of different length, e.g. leap years).  Every
 * later nesosim_run_season advances each member through ITS season; output slots beyond a member's last day are left
 * untouched.  densityType='variable' only.  nesosim_set_forcing returns the context to a single shared season. */
int nesosim_set_forcing_sets(nesosim_ctx *ctx, int n_sets, const double *precip_dev, const double *conc_dev,
                             const double *wind_dev, const double *drift_dev, const int32_t *member_set_host,
                             const int32_t *set_days_host);

/* The season: IC handling of main (NESOSIM.py:604-609; ic_dev = [ny][nx] total depth shared by all members,
 * or [M][ny][nx] when ic_per_member != 0, or NULL for zero depth) when first_step == 0, then steps
 * x = first_step .. first_step+num_steps-1 of `for x in range(numDays-1): calcBudget(...)`
 * (NESOSIM.py:614-639) for every member.  Slot 0 of every non-NULL output is (re)written when
 * first_step == 0 (zeros; IC halves for snowDepths) so callers need not zero-fill.  num_steps < 0 means
 * "to the end" (T-1-first_step).  Asynchronous on `stream`, except that the season-resident path (see
 * nesosim_set_path) synchronises the stream once at the end to read its operand-range flag: that kernel only
 * carries the proven-exact fast divisions, and a season in which some operand left their range (denormal or
 * > 2^624 magnitudes; never on physical data) is transparently redone by the general per-day kernels. */
int nesosim_run_season(nesosim_ctx *ctx, const nesosim_member_params *params_host, const double *ic_dev,
                       int ic_per_member, const nesosim_outputs *out, int first_step, int num_steps,
                       void *stream);

/* One calcBudget call (NESOSIM.py:224-347) with that day's forcing planes given explicitly: advances slot x
 * to slot x+1 of `out` in place for every member.  rho_new is used only when density_clim=1. */
int nesosim_step_day(nesosim_ctx *ctx, int x, const double *conc_dev, const double *precip_dev,
                     const double *drift_dev, const double *wind_dev, double rho_new,
                     const nesosim_member_params *params_host, const nesosim_outputs *out, void *stream);

/* Same as nesosim_run_season but every data pointer is HOST memory (pinned or pageable): copies the forcing
 * to the device, runs the season, copies every non-NULL output back, synchronises.  This is the call
 * bench.py times for the end-to-end figure.  `out_host` strides are in elements like nesosim_outputs.
 * Members are processed in batches so that device and pinned staging memory stay bounded; bytes moved are
 * reported through h2d_bytes / d2h_bytes when non-NULL.  snowAcc and snowOcean do not depend on the member (forcing
 * only, NESOSIM.py:263-270): with one shared forcing a single copy crosses the link and host threads replicate it into
 * every member's slot of the caller's arrays.  When the process has at least 6 host threads to itself
 * (NESOSIM_HOST_THREADS; default: cores / visible GPUs) and at most 60 % of the grid is ocean, the other nine arrays are
 * drained in packed form -- ocean cells, plus the land cells of the first three time slots -- and scattered into the
 * caller's arrays by those threads (NESOSIM_HOST_COMPACT=0/1 overrides); the arrays are the same either way
 * (nesosim_host_drain_info). */
int nesosim_run_season_host(nesosim_ctx *ctx, const double *precip, const double *conc, const double *wind,
                            const double *drift, const double *rho_clim,
                            const nesosim_member_params *params, const double *ic, int ic_per_member,
                            const nesosim_outputs *out_host, int64_t *h2d_bytes, int64_t *d2h_bytes);

/* smooth_snow (NESOSIM.py:170-187) = astropy convolve(arr, 3x3 kernel) with defaults, both branches
 * (plain, and NaN-interpolating when the input sum is NaN).  in/out are distinct [ny][nx] device planes. */
int nesosim_smooth(const double *in_dev, double *out_dev, int ny, int nx, const double weights_host[9],
                   double divisor, void *stream);

/* Per-function entry points (device planes), one per reference function, for known-answer tests. */
/* calcDynamics (NESOSIM.py:189-222): drift [2][ny][nx], depths [2][ny][nx] -> adv [2][ny][nx], div [2][ny][nx] */
int nesosim_op_dynamics(const double *drift_dev, const double *depths_dev, double dx, double deltaT,
                        int ny, int nx, double *adv_dev, double *div_dev, void *stream);
/* calcLeadLoss / calcAtmLoss / calcWindPacking (NESOSIM.py:51-125): five output planes, any may be NULL */
int nesosim_op_wind_terms(const double *h0_dev, const double *wind_dev, const double *conc_dev, int64_t n,
                          const nesosim_member_params *p_host, double deltaT, double rhoFresh, double rhoOld,
                          double *lead_dev, double *atm_dev, double *wp_loss_dev, double *wp_gain_dev,
                          double *wp_net_dev, void *stream);
/* fillMaskAndNaNWithZero (NESOSIM.py:127-139), in place */
int nesosim_op_fill_zero(double *arr_dev, int64_t n, void *stream);
/* fill_nan_no_negative (NESOSIM.py:141-166), in place; mask_dev is uint8 [n] on the device */
int nesosim_op_fill_nan_no_negative(double *arr_dev, const uint8_t *mask_dev, int64_t n,
                                    int negative_to_zero, void *stream);
/* densityCalc (NESOSIM.py:458-473): depths [2][n] -> density [n] */
int nesosim_op_density(const double *depths_dev, const uint8_t *mask_dev, int64_t n, double rhoFresh,
                       double rhoOld, double minSnowD, double *density_dev, void *stream);

/* Final-product diagnostics: the array preparation of OutputSnowModelFinal (utils.py:161-179) on what main hands it
 * (NESOSIM.py:654: snowVol = h0+h1, snowDepth = snowVol/iceConc), fused into one pass and written as the float32
 * fields the NetCDF file stores: every field np.around(x, 4) (= rint(x*1e4)/1e4 in fp64) then cast to float32;
 * snow_volume, snow_depth and snow_density are NaN where iceConc < ice_conc_mask (skipped when ice_conc_mask <= 0),
 * ice_concentration is then NaN where iceConc < 0.15.  Inputs are device arrays of one member: depths [T][2][plane],
 * everything else [T][plane]; any output may be NULL.  96 B -> 24 B per cell-day on the way to the host.
 * For an ensemble pass the members' stacked arrays as num_days = M*T and forcing_days = T: depths/density (and the
 * outputs) then hold M*T days while conc/precip/wind hold T days that every member shares (0 = same as num_days). */
int nesosim_final_products(const double *depths_dev, const double *density_dev, const double *conc_dev,
                           const double *precip_dev, const double *wind_dev, int num_days, int forcing_days, int64_t plane,
                           double ice_conc_mask, float *snow_depth_dev, float *snow_volume_dev,
                           float *snow_density_dev, float *ice_conc_dev, float *precip_out_dev, float *wind_out_dev,
                           void *stream);

/* Calibration driver (the reference has no counterpart): the season of nesosim_run_season with the
 * observation sampling FUSED into the season-resident kernel -- no output array is written at all.
 * nesosim_set_observations registers point observations (day slot 0..T-1, row, col, observed snow depth over ice [m];
 * host arrays, copied): they are compiled against the kernel's cell ownership once (distinct (cell, day) pairs =
 * "sample slots", sorted per owning thread) and stay on the device until replaced.  During nesosim_run_season_misfit the
 * thread that owns an observed cell stores the total depth h0+h1 of the observed days -- 8 bytes per observed cell-day
 * instead of 96 per cell-day, nothing on the day's critical path waits for it -- and a one-CTA-per-member epilogue forms
 * model = (h0+h1)/iceConc at that slot and cell (what main writes as snow depth, NESOSIM.py:654) and
 * misfit_dev[m] = sum over the observations of (model - observed)^2, skipping non-finite differences (land, NaN forcing,
 * zero concentration); count_dev[m] (may be NULL) = the number of observations used.  Sums are formed in a fixed order
 * (bit-reproducible).  Needs the season-resident path (grid up to 96 columns, variable density, one shared forcing);
 * NESOSIM_ERR_ARG otherwise. */
int nesosim_set_observations(nesosim_ctx *ctx, int64_t n_obs, const int32_t *obs_day_host, const int32_t *obs_row_host,
                             const int32_t *obs_col_host, const double *obs_depth_host);
int nesosim_run_season_misfit(nesosim_ctx *ctx, const nesosim_member_params *params_host, const double *ic_dev,
                              int ic_per_member, double *misfit_dev, int64_t *count_dev, void *stream);

/* Asynchronous mode.  By default nesosim_run_season on the season-resident path synchronises `stream` once to read
 * the kernel's operand-range flag (see above).  With nesosim_set_async(ctx, 1) it never synchronises: the flag is copied
 * to pinned host memory behind the kernel and examined by nesosim_sync (which waits for every season enqueued so far on
 * this context, redoes with the general kernels any season whose flag is raised -- never on physical data -- and
 * reports how many) or, without waiting, by later calls.  Contract in this mode: a season's outputs, its IC buffer and
 * the forcing are final / reusable only after nesosim_sync; back-to-back seasons, the next forcing's upload and the
 * previous outputs' download then overlap freely on the caller's streams.  nesosim_set_async(ctx, 0) syncs first. */
int nesosim_set_async(nesosim_ctx *ctx, int on);
int nesosim_sync(nesosim_ctx *ctx, int *seasons_redone);

/* Kernel path of nesosim_run_season: 0 = automatic (default), 1 = general per-day kernel (any grid),
 * 2 = season-resident cluster kernel (grids up to 96x96, variable density, whole season; error otherwise).
 * Both paths produce identical values.  nesosim_last_path reports which one the last season used. */
int nesosim_set_path(nesosim_ctx *ctx, int path);
int nesosim_last_path(const nesosim_ctx *ctx);

/* ---- Row-strip domain decomposition over peer memory (grids too large for one GPU: the 5 km case; the reference
 * has no counterpart -- its grid is one numpy array -- so the contract is that of calcBudget on the whole grid).
 * A context created on rows [lo-2, hi+2) of the full grid (mask and forcing sliced the same way; 2 = radius of
 * np.gradient followed by the 3x3 convolution, NESOSIM.py:204-213,184-185; no ghost rows at the global edges) is one
 * STRIP.  After nesosim_strip_setup + nesosim_strip_connect*, nesosim_run_season on the general path advances the
 * strip with ONE launch per day and nothing in between: the day kernel itself stores the new depths of its first /
 * last two owned rows into the neighbouring strip's mailbox (peer memory over NVLink when the neighbour is another
 * GPU) and raises that strip's flag; the next day's boundary CTAs wait for their own flag and take the ghost rows
 * from the mailbox.  The owned rows of every output come out identical to a single-context run of the whole grid;
 * the ghost rows of the outputs are scratch.  n_members must be 1.  All strips must run the same sequence of
 * nesosim_run_season calls (whole seasons, or the same first_step/num_steps pieces), and every strip's previous
 * season must have completed on its stream before any strip starts the next one (one barrier between seasons).
 * A neighbour that never delivers does not hang the GPU: waits give up after a time-out (default 5 s) and
 * nesosim_strip_status reports it. */
#define NESOSIM_IPC_HANDLE_BYTES 64
int nesosim_strip_setup(nesosim_ctx *ctx, int has_up_neighbour, int has_down_neighbour);
/* the strip's exchange block: as a cudaIpcMemHandle_t (64 bytes, for a neighbour in another process) ... */
int nesosim_strip_export(nesosim_ctx *ctx, void *handle64);
/* ... or as a device pointer (for a neighbour in the same process; other device: peer access must be enabled) */
int nesosim_strip_block(nesosim_ctx *ctx, void **block_dev, int64_t *bytes);
/* attach the neighbours' blocks (NULL where there is no neighbour) */
int nesosim_strip_connect(nesosim_ctx *ctx, const void *up_handle64, const void *down_handle64);
int nesosim_strip_connect_local(nesosim_ctx *ctx, void *up_block_dev, void *down_block_dev);
/* *timed_out = 1 if any wait for a neighbour gave up since nesosim_strip_setup (synchronous device read) */
int nesosim_strip_status(nesosim_ctx *ctx, int *timed_out);
int nesosim_strip_set_timeout(nesosim_ctx *ctx, double seconds);

/* Diagnostics for the exact constant-division path used for /dx, /(2.*dx), /rho and /kernel.sum()
 * (cell_math.cuh div_const): whether the 3-operation path was proven exact for divisor c, and a host
 * replica of the device routine (same branches, std::fma) so CPU tests can compare it with x / c. */
int    nesosim_const_div_is_fast(double c);
double nesosim_const_div_eval_host(double x, double c);

/* Seasons the season-resident path had to hand back to the general kernels (see nesosim_run_season). */
int64_t nesosim_rerun_count(const nesosim_ctx *ctx);

/* How the last nesosim_run_season_host drained its results: *compacted = 1 if only the ocean cells (and the land cells
 * of the first three time slots) of the member-dependent arrays crossed the link and host threads scattered them into
 * the caller's arrays; *full_chunks = (member, array) blocks (since creation) for which that was given up because their land
 * cells were not constant in time, and which were copied in full instead.  The caller's arrays are the same either way
 * (the reference's genEmptyArrays contract, NESOSIM.py:350-376). */
int nesosim_host_drain_info(const nesosim_ctx *ctx, int *compacted, int64_t *full_chunks);

/* The compacted drain shares the work between the host threads and the copy engine as it goes: whenever every slot of
 * the pinned ring is taken (the threads are the bottleneck), the link copies whole (member, array) blocks of the same
 * batch straight into the caller's arrays instead.  *packed / *plain = the blocks of the last call that went either
 * way (both 0 after a plain drain). */
int nesosim_host_drain_blocks(const nesosim_ctx *ctx, int64_t *packed, int64_t *plain);

/* Host half of that drain, on its own: one member's packed block of one array -- [planes_per_slot*num_days][n_ocean]
 * ocean values (cells with mask 1..10, ascending), then [planes_per_slot*min(3,num_days)][n_land] land values of the
 * first slots -- scattered into the full planes dst[planes_per_slot*num_days][plane]; later slots repeat the land cells
 * of the third.  No device involved. */
int nesosim_unpack_member_array(const uint8_t *mask, int64_t plane, int planes_per_slot, int num_days,
                                const double *packed, double *dst);

/* Device time of the season-resident kernel's launches so far (CUDA events on the launching stream) and their
 * number: what bench.py divides the algorithmic bytes per launch by for its roofline figure. */
int nesosim_season_kernel_time(const nesosim_ctx *ctx, double *total_ms, int64_t *launches);

/* Number of kernel launches this context has issued since creation (bench.py reports it). */
int64_t nesosim_launch_count(const nesosim_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* NESOSIM_B200_H */
