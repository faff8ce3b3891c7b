// drain_kernels.cuh -- compacted device->host drain of nesosim_run_season_host (the e2e leg).
//
// The reference's output contract is full (T,ny,nx) arrays (genEmptyArrays, NESOSIM.py:350-376), but more than half
// of every plane is land (region mask > 10 or < 1; fill_nan_no_negative, NESOSIM.py:141-166) and a land cell of a
// member-dependent array does not change after the first steps: its depths are NaN from step 1 on, so every term fed by
// them is NaN (or stays what it was when its switch is off).  The link to the host is what bounds the end-to-end call,
// so only the ocean cells of every plane cross it, plus the land cells of the first DRAIN_HEAD time slots; the host
// scatters them into the caller's arrays and repeats slot DRAIN_HEAD-1's land cells for the later slots.  That the land
// cells really are bit-constant from there on is not assumed: the pack kernel compares them and counts the differences
// per member and array, and a flagged member's array is copied in full instead.
#pragma once
#include <cstdint>

namespace nesosim {

constexpr int DRAIN_HEAD = 3;        // time slots whose land cells are shipped as they are
constexpr int DRAIN_MAX_ARRAYS = 10; // member-dependent arrays of nesosim_outputs (all but snowAcc / snowOcean)

struct PackArgs {
    const double *src[DRAIN_MAX_ARRAYS];   // member 0 of the chunk, slot 0 of each array (device)
    long long mstride[DRAIN_MAX_ARRAYS];   // elements between two members of src[a]
    long long rec_off[DRAIN_MAX_ARRAYS];   // where array a starts inside a member's packed record (elements)
    int pps[DRAIN_MAX_ARRAYS];             // planes per time slot (2 for snowDepths, 1 otherwise)
    int plane0[DRAIN_MAX_ARRAYS + 1];      // prefix sum of pps[a] * T: blockIdx.x -> (array, plane)
    int n_arrays, T, head;                 // head = min(DRAIN_HEAD, T)
    int plane, n_ocean, n_land;
    const int *ocean_idx, *land_idx;       // cell indices, ascending
    double *dst;                           // packed records of the chunk's members, member-major
    long long rec_elems;                   // elements per member record
    unsigned long long *flag;              // [member][array]: += 1 for every land cell that is not bit-constant after the head slots
};

// Packed record of one member: for every array a, [pps*T planes][n_ocean] ocean values followed by
// [head*pps planes][n_land] land values of the first `head` slots.
__global__ void __launch_bounds__(256) pack_ocean_kernel(const PackArgs a) {
    int arr = 0;
    while (arr + 1 < a.n_arrays && (int)blockIdx.x >= a.plane0[arr + 1]) ++arr;
    const int q = (int)blockIdx.x - a.plane0[arr];      // plane of this array: slot * pps + layer
    const int pps = a.pps[arr];
    const int slot = q / pps, layer = q - slot * pps;
    const int m = blockIdx.y;
    const double *src = a.src[arr] + (long long)m * a.mstride[arr] + (long long)q * a.plane;
    double *rec = a.dst + (long long)m * a.rec_elems + a.rec_off[arr];
    double *oc = rec + (long long)q * a.n_ocean;
    for (int k = threadIdx.x; k < a.n_ocean; k += 256) oc[k] = src[__ldg(a.ocean_idx + k)];
    if (slot < a.head) {
        double *ld = rec + (long long)pps * a.T * a.n_ocean + (long long)q * a.n_land;
        for (int l = threadIdx.x; l < a.n_land; l += 256) ld[l] = src[__ldg(a.land_idx + l)];
    } else {
        const double *ref = a.src[arr] + (long long)m * a.mstride[arr] + (long long)((a.head - 1) * pps + layer) * a.plane;
        unsigned bad = 0;
        for (int l = threadIdx.x; l < a.n_land; l += 256) {
            const int c = __ldg(a.land_idx + l);
            bad += __double_as_longlong(src[c]) != __double_as_longlong(ref[c]);
        }
        if (bad) atomicAdd(a.flag + (long long)m * a.n_arrays + arr, (unsigned long long)bad);
    }
}

}  // namespace nesosim
