// Hand-written sm_100a kernels of the ganon-classify hot path.
//
//   K1  k_count_newlines / k_line_starts / k_fastq_records   FASTQ block -> record table   (reader GC.cpp:1220-1287,
//                                                             seqan3 format_fastq.hpp:105-267)
//   K2  k_minimisers        warp per read: canonical minimiser hashes (minimiser_hash.hpp:76-108, minimiser.hpp:398-472,
//                           kmer_hash.hpp:618-640, dna4.hpp:166-205)
//   K3  k_ibf_count         warp per (read, 4096-bin chunk): h row gathers with 128-bit loads, AND in registers,
//                           bit-sliced (carry-save) per-bin counters, cutoff + sparse emission
//                           (bulk_count IBF.hpp:1027-1042, bulk_contains 639-664, hash_and_fit 173-187,
//                           counting_vector += 926-953; select_matches GC.cpp:504-541)
//   aux k_fill_random / k_emplace   build-side helpers (IBF.hpp:271-286)
//
// These are HBM-bound integer kernels: no tensor-core work exists on this path.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include "gnb_internal.h"
#include "k2_thread.cuh"

namespace gnb
{

// =====================================================================================================================
// K2 -- minimisers
// =====================================================================================================================
namespace
{

constexpr int K2_WARPS = 8;   // warps (= reads in flight) per CTA
constexpr int K2_TILE  = 128; // windows per tile (up to 4 consecutive windows per lane)

// seqan3 dna4 char_to_rank incl. IUPAC conversion (dna4.hpp:166-205): everything not listed maps to 0 ('A').
__device__ __forceinline__ uint32_t dna4_rank(uint8_t c)
{
    c |= 0x20; // lower case
    uint32_t r = 0;
    r = (c == 'c' || c == 'y' || c == 's' || c == 'b') ? 1u : r;
    r = (c == 'g' || c == 'k') ? 2u : r;
    r = (c == 't' || c == 'u') ? 3u : r;
    return r;
}

// Rightmost minimum of sv[0..W) (std::ranges::min_element with less_equal, minimiser.hpp:433-435); dup = the minimum
// occurs more than once.
__device__ __forceinline__ uint32_t rightmost_min(const uint64_t *sv, uint32_t W, uint64_t &best, bool &dup)
{
    uint32_t idx = 0;
    best = sv[0];
    bool d = false;
    for (uint32_t x = 1; x < W; ++x)
    {
        const uint64_t y = sv[x];
        if (y < best)
        {
            best = y;
            idx  = x;
            d    = false;
        }
        else if (y == best)
        {
            idx = x;
            d   = true;
        }
    }
    dup |= d;
    return idx;
}

// bases [b0, b0+nb) of the mate -> ranks in shared memory, then canonical k-mer values v[0..nv) (value i = k-mer at
// base b0 + i); each lane rolls over a contiguous run (kmer_hash.hpp:618-640, minimiser_hash.hpp:91-107)
__device__ __forceinline__ void load_values(const uint8_t *__restrict__ seq, uint32_t nv, uint32_t k, uint64_t seed, uint64_t kmask,
                                            uint64_t *sv, uint8_t *sb, uint32_t lane)
{
    const uint32_t nb = nv + k - 1;
    __syncwarp();
    for (uint32_t i = lane; i < nb; i += 32)
        sb[i] = (uint8_t)dna4_rank(seq[i]);
    __syncwarp();
    const uint32_t seg = (nv + 31) >> 5;
    const uint32_t i0 = lane * seg, i1 = min(nv, i0 + seg);
    if (i0 < i1)
    {
        uint64_t f = 0, r = 0;
        for (uint32_t j = 0; j < k; ++j)
        {
            const uint64_t b = sb[i0 + j];
            f = (f << 2) | b;                          // first base most significant
            r = (r >> 2) | ((3 - b) << (2 * (k - 1))); // reverse complement, complement = 3 - rank (dna4.hpp:95-98)
        }
        sv[i0] = min(f ^ seed, r ^ seed);
        for (uint32_t i = i0 + 1; i < i1; ++i)
        {
            const uint64_t b = sb[i + k - 1];
            f = ((f << 2) | b) & kmask;
            r = (r >> 2) | ((3 - b) << (2 * (k - 1)));
            sv[i] = min(f ^ seed, r ^ seed);
        }
    }
    __syncwarp();
}

// ---- fast path (k <= 31): packed 2-bit bases, doubling-tree window minimum -----------------------------------------
// rank table of dna4_rank packed 2 bits per letter 'a'..'z'
__device__ __forceinline__ uint32_t dna4_rank_fast(uint8_t c)
{
    // c:1 y:1 s:1 b:1 | g:2 k:2 | t:3 u:3 ; everything else 0
    constexpr uint64_t T = (1ull << (2 * ('c' - 'a'))) | (1ull << (2 * ('y' - 'a'))) | (1ull << (2 * ('s' - 'a'))) | (1ull << (2 * ('b' - 'a'))) |
                           (2ull << (2 * ('g' - 'a'))) | (2ull << (2 * ('k' - 'a'))) | (3ull << (2 * ('t' - 'a'))) | (3ull << (2 * ('u' - 'a')));
    const uint32_t i = (uint32_t)(c | 0x20) - 'a';
    const uint32_t lo = (uint32_t)T, hi = (uint32_t)(T >> 32);
    const uint32_t w = i < 16 ? lo : hi;
    return i < 26 ? (w >> ((i & 15) * 2)) & 3u : 0u;
}

// keys: (canonical k-mer value << 1) | dup-flag; comb = minimum by value, equal values set the flag.
__device__ __forceinline__ uint64_t key_comb(uint64_t a, uint64_t b)
{
    const uint64_t m = a < b ? a : b;
    return ((a ^ b) < 2) ? (m | 1ull) : m;
}

// bases [0, nv+k-1) of `seq` -> 2-bit packed big-endian words (REDUX.OR over the warp) -> keys lev0[0..nv)
__device__ __forceinline__ void load_keys(const uint8_t *__restrict__ seq, uint32_t nv, uint32_t k, uint64_t seed, uint64_t kmask, uint64_t *sp,
                                          uint64_t *lev0, uint32_t lane)
{
    const uint32_t nb = nv + k - 1;
    const uint32_t ng = (nb + 31) >> 5;
    __syncwarp();
    for (uint32_t t = 0; t < ng; ++t)
    {
        const uint32_t i  = t * 32 + lane;
        const uint32_t r  = i < nb ? dna4_rank_fast(seq[i]) : 0u;
        const uint32_t hi = __reduce_or_sync(0xffffffffu, lane < 16 ? r << (30 - 2 * lane) : 0u);
        const uint32_t lo = __reduce_or_sync(0xffffffffu, lane >= 16 ? r << (62 - 2 * lane) : 0u);
        if (lane == 0)
            sp[t] = ((uint64_t)hi << 32) | lo;
    }
    if (lane == 0)
        sp[ng] = 0;
    __syncwarp();
    const uint32_t sh = 64 - 2 * k;
    for (uint32_t i = lane; i < nv; i += 32)
    {
        const uint32_t j = i >> 5, o = (i & 31) * 2;
        const uint64_t a = sp[j], b = sp[j + 1];
        const uint64_t x = (a << o) | ((b >> 1) >> (63 - o)); // k-mer in the top 2k bits, first base most significant
        const uint64_t f = x >> sh;
        uint64_t       z = __brevll(~x);                      // complement, reverse: bases in reverse order, bit pairs swapped
        z = ((z >> 1) & 0x5555555555555555ull) | ((z & 0x5555555555555555ull) << 1);
        const uint64_t rc = z & kmask;
        const uint64_t fs = f ^ seed, rs = rc ^ seed;
        lev0[i] = (fs < rs ? fs : rs) << 1;
    }
    __syncwarp();
}

// ---- fastest path (k <= 26, 4 <= W <= 255): position-tagged keys, a run of C consecutive windows per lane ------------
// Keys (value << 9) | tag are totally ordered: tag = 511 - p makes the minimum the RIGHTMOST minimal value (R), tag = p
// the leftmost (L); R != L in some window means a duplicated minimum -> the caller falls back to the exact serial walk.
// A lane owns windows [lane*C, lane*C + C): they share the values at offsets [C-1, W), so the lane takes the minimum of
// that shared part once, suffix minima over the C-1 values in front and prefix minima over the C-1 values behind, and
// every window is the minimum of three terms: (W + 3C - 4) 64-bit minima for C windows instead of C*(W-1).
// Values are stored residue-major (position p at (p % C) * pitch + p / C) so that the lanes' reads are conflict-free.
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

template <bool WRITE, int C>
__device__ __forceinline__ bool window_runs(const uint64_t *sv, uint32_t pitch, uint32_t nv, uint32_t nt, uint32_t W, uint32_t t0, uint32_t lane,
                                            uint32_t &prev_r, uint32_t &emitted, uint64_t *__restrict__ out)
{
    constexpr uint64_t BIG = 0x3ffffffffffffe00ull; // larger than any key, tag bits clear
    const uint32_t i = lane * C;                    // first window (tile-local) of this lane
    auto key = [&](uint32_t t, uint64_t &kr, uint64_t &kl) {
        const uint32_t p = i + t;
        const uint64_t u = p < nv ? sv[(p % C) * pitch + p / C] : BIG;
        kr = u | (uint64_t)(511u - (p & 511u));
        kl = u | (uint64_t)(p & 511u);
    };
    // shared part
    uint64_t mR, mL;
    key(C - 1, mR, mL);
    for (uint32_t t = C; t < W; ++t)
    {
        uint64_t a, b;
        key(t, a, b);
        mR = umin64(mR, a);
        mL = umin64(mL, b);
    }
    // suffix minima of the values in front: sufR[j] = min over offsets [j, C-1)
    uint64_t sufR[C], sufL[C];
    sufR[C - 1] = sufL[C - 1] = ~0ull;
#pragma unroll
    for (int j = C - 2; j >= 0; --j)
    {
        uint64_t a, b;
        key((uint32_t)j, a, b);
        sufR[j] = umin64(sufR[j + 1], a);
        sufL[j] = umin64(sufL[j + 1], b);
    }
    // windows in order, extending the prefix minimum of the values behind
    uint64_t preR = ~0ull, preL = ~0ull;
    uint32_t R[C];
    uint64_t val[C];
    bool     tie = false;
#pragma unroll
    for (int j = 0; j < C; ++j)
    {
        if (j > 0)
        {
            uint64_t a, b;
            key(W + (uint32_t)j - 1, a, b);
            preR = umin64(preR, a);
            preL = umin64(preL, b);
        }
        const uint64_t wr = umin64(umin64(sufR[j], mR), preR), wl = umin64(umin64(sufL[j], mL), preL);
        const bool     on = i + j < nt;
        R[j]   = 511u - (uint32_t)(wr & 511u);
        val[j] = wr >> 9;
        tie |= on && R[j] != (uint32_t)(wl & 511u);
    }
    if (__any_sync(0xffffffffu, tie))
        return false;
    // emissions: window j emits iff its minimum sits elsewhere than the previous window's
    const uint32_t my_last = i < nt ? R[min((uint32_t)C, nt - i) - 1] + t0 : 0u;
    uint32_t       left    = __shfl_up_sync(0xffffffffu, my_last, 1);
    if (lane == 0)
        left = prev_r;
    const uint32_t last_lane = (nt - 1) / C;
    prev_r = __shfl_sync(0xffffffffu, my_last, last_lane);
    uint32_t emask = 0;
#pragma unroll
    for (int j = 0; j < C; ++j)
    {
        const uint32_t rg = R[j] + t0;
        if (i + j < nt && rg != left)
            emask |= 1u << j;
        left = rg;
    }
    const uint32_t mine = __popc(emask);
    uint32_t       incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += y;
    }
    if (WRITE)
    {
        uint32_t pos = emitted + incl - mine;
#pragma unroll
        for (int j = 0; j < C; ++j)
            if ((emask >> j) & 1u)
                out[pos++] = val[j];
    }
    emitted += __shfl_sync(0xffffffffu, incl, 31);
    return true;
}

template <bool WRITE>
__device__ __forceinline__ bool minimisers_runs(const uint8_t *__restrict__ seq, uint32_t nwin, uint32_t W, uint32_t k, uint64_t seed,
                                                uint64_t kmask, uint64_t *__restrict__ out, uint64_t *sp, uint64_t *sv, uint32_t lane,
                                                uint32_t &emitted_out)
{
    const uint32_t sh   = 64 - 2 * k;
    const uint32_t span = 32 - k + 1; // k-mers available from one 32-base word pair
    uint32_t emitted = 0;
    uint32_t prev_r  = 0xffffffffu; // global position of the previous window's minimum
    for (uint32_t t0 = 0; t0 < nwin; t0 += K2_TILE)
    {
        const uint32_t nt = min((uint32_t)K2_TILE, nwin - t0);
        const uint32_t nv = nt + W - 1;
        const uint32_t nb = nv + k - 1;
        const uint32_t ng = (nb + 31) >> 5;
        const uint32_t C  = (nt + 31) >> 5;      // windows per lane, 1..8
        const uint32_t pitch = (nv + C - 1) / C; // values per residue class
        const uint8_t *sq = seq + t0;
        __syncwarp();
        // bases -> 2-bit packed words; the byte loads of six words are issued together (memory-level parallelism)
        for (uint32_t tb = 0; tb < ng; tb += 6)
        {
            uint8_t c[6];
#pragma unroll
            for (int q = 0; q < 6; ++q)
            {
                const uint32_t i = (tb + q) * 32 + lane;
                c[q] = i < nb ? sq[i] : (uint8_t)'A';
            }
#pragma unroll
            for (int q = 0; q < 6; ++q)
                if (tb + q < ng)
                {
                    const uint32_t r  = dna4_rank_fast(c[q]);
                    const uint32_t hi = __reduce_or_sync(0xffffffffu, lane < 16 ? r << (30 - 2 * lane) : 0u);
                    const uint32_t lo = __reduce_or_sync(0xffffffffu, lane >= 16 ? r << (62 - 2 * lane) : 0u);
                    if (lane == 0)
                        sp[tb + q] = ((uint64_t)hi << 32) | lo;
                }
        }
        if (lane == 0)
            sp[ng] = 0;
        __syncwarp();
        // canonical values (<< 9) of a contiguous run of positions per lane
        {
            const uint32_t rl = (nv + 31) >> 5;
            const uint32_t i0 = lane * rl, i1 = min(nv, i0 + rl);
            uint64_t x = 0, z = 0;
            for (uint32_t i = i0, d = span; i < i1; ++i, ++d)
            {
                if (d >= span)
                { // (re)load the 32 bases starting at position i
                    const uint32_t j = i >> 5, o = (i & 31) * 2;
                    const uint64_t a = sp[j], b = sp[j + 1];
                    x = (a << o) | ((b >> 1) >> (63 - o));
                    z = __brevll(~x);
                    z = ((z >> 1) & 0x5555555555555555ull) | ((z & 0x5555555555555555ull) << 1); // comp(base t) at bits 2t+1..2t
                    d = 0;
                }
                const uint64_t f  = ((x << (2 * d)) >> sh) ^ seed;
                const uint64_t rc = ((z >> (2 * d)) & kmask) ^ seed;
                sv[(i % C) * pitch + i / C] = umin64(f, rc) << 9;
            }
        }
        __syncwarp();
        bool ok;
        switch (C)
        {
        case 1: ok = window_runs<WRITE, 1>(sv, pitch, nv, nt, W, t0, lane, prev_r, emitted, out); break;
        case 2: ok = window_runs<WRITE, 2>(sv, pitch, nv, nt, W, t0, lane, prev_r, emitted, out); break;
        case 3: ok = window_runs<WRITE, 3>(sv, pitch, nv, nt, W, t0, lane, prev_r, emitted, out); break;
        default: ok = window_runs<WRITE, 4>(sv, pitch, nv, nt, W, t0, lane, prev_r, emitted, out); break;
        }
        if (!ok)
            return false; // duplicated minimum inside a window
    }
    emitted_out = emitted;
    return true;
}

// Exact for every input: windows with a duplicated minimum make the caller fall back to the serial walk.
// Per tile: keys -> min-with-dup-flag over 2,4,8,... consecutive values (in shared memory) -> window minimum as the
// fold over the binary digits of W -> window i emits iff the previous minimum left the window (v[i-1] == m(i-1)) or
// the newcomer is smaller (v[i+W-1] < m(i-1)).
template <bool WRITE>
__device__ __forceinline__ bool minimisers_fast(const uint8_t *__restrict__ seq, uint32_t nwin, uint32_t W, uint32_t k, uint64_t seed,
                                                uint64_t kmask, uint64_t *__restrict__ out, uint64_t *sp, uint64_t *lev, uint32_t nv_cap,
                                                uint32_t lane, uint32_t &emitted_out)
{
    const uint32_t lmax = 31 - __clz(W); // levels 0..lmax
    uint32_t emitted = 0;
    uint64_t prev    = 0; // window minimum key of the window before the current group (carried in every lane)
    for (uint32_t t0 = 0; t0 < nwin; t0 += K2_TILE)
    {
        const uint32_t nt   = min((uint32_t)K2_TILE, nwin - t0);
        const uint32_t base = t0 ? 1u : 0u; // one extra value in front: v[i-1] of the tile's first window
        const uint32_t nv   = nt + W - 1 + base;
        load_keys(seq + t0 - base, nv, k, seed, kmask, sp, lev, lane);
        for (uint32_t l = 1; l <= lmax; ++l)
        {
            const uint64_t *src = lev + (size_t)(l - 1) * nv_cap;
            uint64_t       *dst = lev + (size_t)l * nv_cap;
            const uint32_t  h   = 1u << (l - 1);
            for (uint32_t i = lane; i + 2 * h <= nv; i += 32)
                dst[i] = key_comb(src[i], src[i + h]);
            __syncwarp();
        }
        for (uint32_t g = 0; g < nt; g += 32)
        {
            const uint32_t i  = g + lane;
            const bool     on = i < nt;
            uint64_t       acc = ~0ull;
            if (on)
            {
                uint32_t pos = i + base;
                for (int l = (int)lmax; l >= 0; --l)
                    if ((W >> l) & 1u)
                    {
                        const uint64_t term = lev[(size_t)l * nv_cap + pos];
                        acc = acc == ~0ull ? term : key_comb(acc, term);
                        pos += 1u << l;
                    }
            }
            if (__any_sync(0xffffffffu, on && (acc & 1ull)))
                return false; // tie inside a window
            uint64_t left = __shfl_up_sync(0xffffffffu, acc, 1);
            if (lane == 0)
                left = prev;
            const uint32_t last = min(31u, nt - 1 - g);
            prev = __shfl_sync(0xffffffffu, acc, last);
            bool emit = false;
            if (on)
            {
                if (t0 + i == 0)
                    emit = true;
                else
                {
                    const uint64_t gone = lev[i + base - 1], come = lev[i + base + W - 1]; // keys (value << 1)
                    emit = (gone >> 1) == (left >> 1) || (come >> 1) < (left >> 1);
                }
            }
            const uint32_t mask = __ballot_sync(0xffffffffu, emit);
            if (WRITE && emit)
                out[emitted + __popc(mask & ((1u << lane) - 1))] = acc >> 1;
            emitted += __popc(mask);
        }
    }
    emitted_out = emitted;
    return true;
}

// One mate.  Returns the number of minimisers emitted (warp-uniform).
//
// The reference's window (minimiser.hpp:444-472) is a sticky state machine: the tracked minimiser position `mp` only
// moves when a strictly smaller value enters, or when it leaves the window (then the RIGHTMOST minimum of the new
// window is taken and emitted even if the value repeats).  Whenever every window has a unique minimum, mp(i) is simply
// that minimum R(i), and window i emits iff R(i) != R(i-1) (or i == 0).  So consecutive lanes take consecutive windows,
// each scans its W values for the minimum (uniform work, conflict-free shared-memory reads), neighbours exchange R by
// shuffle and the emissions are compacted in order with a ballot.  A duplicated minimum in any window raises `tie`,
// and the whole mate is redone by the exact serial walk below -- homopolymers, short-period repeats etc. stay
// bit-exact.
template <bool WRITE>
__device__ uint32_t minimisers_of_mate(const uint8_t *__restrict__ seq, uint32_t L, uint32_t k, uint32_t w, uint64_t seed,
                                       uint64_t *__restrict__ out, uint64_t *sv, uint8_t *sb, uint64_t *sp, uint32_t nv_cap, uint32_t lane)
{
    const uint32_t nk   = L - k + 1;
    const uint32_t W    = min(w - k + 1, nk);
    const uint32_t nwin = nk - W + 1;
    const uint64_t kmask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);

    uint32_t emitted = 0;
    bool     tie     = false;
    if (k <= 26 && W >= 4)
    {
        if (minimisers_runs<WRITE>(seq, nwin, W, k, seed, kmask, out, sp, sv, lane, emitted))
            return emitted;
        tie     = true;
        emitted = 0;
    }
    else if (k <= 31)
    {
        if (minimisers_fast<WRITE>(seq, nwin, W, k, seed, kmask, out, sp, sv, nv_cap, lane, emitted))
            return emitted;
        tie     = true; // a window with a duplicated minimum: exact serial walk below
        emitted = 0;
    }
    uint32_t prev_r  = 0xffffffffu; // R of the window before the current group of 32 (lane 31's, carried)
    for (uint32_t t0 = 0; k > 31 && t0 < nwin && !tie; t0 += K2_TILE)
    {
        const uint32_t nt = min((uint32_t)K2_TILE, nwin - t0);
        const uint32_t nv = nt + W - 1;
        load_values(seq + t0, nv, k, seed, kmask, sv, sb, lane);
        for (uint32_t g = 0; g < nt; g += 32)
        {
            const uint32_t i  = g + lane; // window [t0+i, t0+i+W-1]
            const bool     on = i < nt;
            uint64_t best = 0;
            bool     dup  = false;
            uint32_t r    = 0;
            if (on)
                r = t0 + i + rightmost_min(sv + i, W, best, dup);
            uint32_t left = __shfl_up_sync(0xffffffffu, r, 1);
            if (lane == 0)
                left = prev_r;
            prev_r = __shfl_sync(0xffffffffu, r, 31);
            const bool     emit = on && r != left; // prev_r starts at an impossible position: window 0 always emits
            const uint32_t mask = __ballot_sync(0xffffffffu, emit);
            if (__any_sync(0xffffffffu, dup))
            {
                tie = true;
                break;
            }
            if (WRITE && emit)
                out[emitted + __popc(mask & ((1u << lane) - 1))] = best;
            emitted += __popc(mask);
        }
    }
    if (!tie)
        return emitted;

    // ---- exact serial walk (ties present) ----
    emitted = 0;
    uint32_t mp = 0;
    uint64_t cur = 0;
    for (uint32_t t0 = 0; t0 < nwin; t0 += K2_TILE)
    {
        const uint32_t nt = min((uint32_t)K2_TILE, nwin - t0);
        const uint32_t nv = nt + W - 1;
        load_values(seq + t0, nv, k, seed, kmask, sv, sb, lane);
        if (lane == 0)
        {
            for (uint32_t i = t0; i < t0 + nt; ++i)
            {
                bool e = false, dummy = false;
                if (i == 0 || mp < i)
                {
                    mp = i + rightmost_min(sv + (i - t0), W, cur, dummy);
                    e  = true;
                }
                else
                {
                    const uint64_t x = sv[i + W - 1 - t0];
                    if (x < cur)
                    {
                        cur = x;
                        mp  = i + W - 1;
                        e   = true;
                    }
                }
                if (e)
                {
                    if (WRITE)
                        out[emitted] = cur;
                    ++emitted;
                }
            }
        }
    }
    __syncwarp();
    return __shfl_sync(0xffffffffu, emitted, 0);
}

// seqan3 dna15 legality as enforced by the readers (dna15.hpp:95, nucleotide_base.hpp:147-168)
__device__ __forceinline__ bool dna15_valid(uint8_t c)
{
    c |= 0x20;
    // a b c d g h k m n r s t u v w y
    const uint32_t ok = (1u << ('a' - 'a')) | (1u << ('b' - 'a')) | (1u << ('c' - 'a')) | (1u << ('d' - 'a')) | (1u << ('g' - 'a')) |
                        (1u << ('h' - 'a')) | (1u << ('k' - 'a')) | (1u << ('m' - 'a')) | (1u << ('n' - 'a')) | (1u << ('r' - 'a')) |
                        (1u << ('s' - 'a')) | (1u << ('t' - 'a')) | (1u << ('u' - 'a')) | (1u << ('v' - 'a')) | (1u << ('w' - 'a')) |
                        (1u << ('y' - 'a'));
    const uint32_t i = (uint32_t)c - 'a';
    return i < 26 && ((ok >> i) & 1u);
}

template <int KMODE> // 0: count only, 1: write at hash_off, 2: write at hash_off and count (single pass)
__global__ void __launch_bounds__(K2_WARPS * 32, 4)
    k_minimisers(const uint8_t *__restrict__ blk1, const uint32_t *__restrict__ off1, const uint32_t *__restrict__ len1,
                 const uint8_t *__restrict__ blk2, const uint32_t *__restrict__ off2, const uint32_t *__restrict__ len2,
                 uint32_t n_reads, uint32_t k, uint32_t w, uint32_t nv_cap, uint32_t nb_cap, uint32_t *__restrict__ counts,
                 const uint64_t *__restrict__ hash_off, uint64_t *__restrict__ hashes, uint32_t *__restrict__ max_count,
                 unsigned long long *__restrict__ sum_count)
{
    constexpr bool WRITE = KMODE != 0;
    uint32_t warp_max = 0;
    uint64_t warp_sum = 0;
    extern __shared__ __align__(16) uint8_t k2_smem[];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // per warp: n_lev key arrays of nv_cap (level 0 doubles as the value array of the generic path), packed words, ranks
    const uint32_t n_lev  = (k <= 26 && w - k + 1 >= 4) ? 1u : 32 - __clz(w - k + 1);
    const uint32_t sp_cap = (nb_cap >> 5) + 2;
    const size_t   per_w  = ((size_t)n_lev * nv_cap + sp_cap) * 8 + nb_cap;
    uint8_t  *wbase = k2_smem + (size_t)wib * per_w;
    uint64_t *sv = reinterpret_cast<uint64_t *>(wbase);
    uint64_t *sp = sv + (size_t)n_lev * nv_cap;
    uint8_t  *sb = reinterpret_cast<uint8_t *>(sp + sp_cap);
    const uint64_t seed = kMinimiserSeed >> (64 - 2 * k);
    const uint32_t wpc = blockDim.x >> 5; // warps per CTA (chosen at launch from the shared-memory need)
    for (uint32_t read = blockIdx.x * wpc + wib; read < n_reads; read += gridDim.x * wpc)
    {
        uint32_t       total = 0;
        const uint32_t L1    = len1[read];
        if (L1 >= w && L1 >= k) // GC.cpp:690: reads shorter than the window are skipped entirely
        {
            uint64_t *out = WRITE ? hashes + hash_off[read] : nullptr;
            total = minimisers_of_mate<WRITE>(blk1 + off1[read], L1, k, w, seed, out, sv, sb, sp, nv_cap, lane);
            if (blk2 != nullptr)
            {
                const uint32_t L2 = len2[read];
                if (L2 >= w && L2 >= k) // GC.cpp:695
                    total += minimisers_of_mate<WRITE>(blk2 + off2[read], L2, k, w, seed, WRITE ? out + total : nullptr, sv, sb, sp, nv_cap, lane);
            }
        }
        if (KMODE != 1 && lane == 0)
            counts[read] = total;
        warp_max = max(warp_max, total);
        warp_sum += total;
    }
    if (KMODE != 1 && lane == 0 && warp_max)
    {
        if (max_count != nullptr)
            atomicMax(max_count, warp_max);
        if (sum_count != nullptr)
            atomicAdd(sum_count, (unsigned long long)warp_sum);
    }
}

// ---- K2t: one thread per read (k2_thread.cuh; the default, GANON_B200_K2=warp switches it off) -------------------------------------------------
// Written after the GPU budget of round 1 was spent: the per-thread core is checked against the oracle on the CPU
// (tests/test_k2t_cpu.py), the kernel has not run on hardware yet, so it is off unless the switch is set.
template <int KMODE> // as k_minimisers
__global__ void __launch_bounds__(k2t::kThreads)
    k_minimisers_thread(const uint8_t *__restrict__ blk1, const uint32_t *__restrict__ off1, const uint32_t *__restrict__ len1, const uint8_t *__restrict__ blk2,
                        const uint32_t *__restrict__ off2, const uint32_t *__restrict__ len2, uint32_t n_reads, uint32_t k, uint32_t w, uint32_t *__restrict__ counts,
                        const uint64_t *__restrict__ hash_off, uint64_t *__restrict__ hashes, uint32_t *__restrict__ max_count, unsigned long long *__restrict__ sum_count,
                        const uint8_t *__restrict__ only)
{
    extern __shared__ __align__(16) uint64_t k2t_smem[]; // [256] character table, then the rings [W][kThreads]: slot-major, a warp's lanes side by side
    k2t::LutEntry *lut = reinterpret_cast<k2t::LutEntry *>(k2t_smem);
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint64_t seed = kMinimiserSeed >> (64 - 2 * k);
    const uint64_t mask = (1ull << (2 * k)) - 1; // k <= 29
    const uint32_t lut_s  = (uint32_t)__cvta_generic_to_shared(k2t_smem);
    const uint32_t ring_s = lut_s + k2t::kLutBytes + tid * 8;
    constexpr uint32_t kStride = k2t::kThreads * 8;
    for (uint32_t c = tid; c < 256; c += k2t::kThreads)
        lut[c] = k2t::lut_entry(c, k);
    __syncthreads();
    uint32_t           my_max = 0;
    unsigned long long my_sum = 0;
    for (uint32_t first = blockIdx.x * k2t::kThreads; first < n_reads; first += gridDim.x * k2t::kThreads)
    {
        const uint32_t read = first + tid;
        if (read >= n_reads)
            continue;
        // only: the segment kernels did the batch, this pass walks the reads they flagged from the start (and sums up the counts)
        const uint32_t total = only != nullptr && !only[read]
                                   ? counts[read]
                                   : k2t::read_pair<KMODE>(read, blk1, off1, len1, blk2, off2, len2, k, w, seed, mask, lut_s, ring_s, kStride, counts, hash_off, hashes);
        my_max = max(my_max, total);
        my_sum += total;
    }
    if (KMODE != 1)
    {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
        {
            my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, d));
            my_sum += __shfl_xor_sync(0xffffffffu, my_sum, d);
        }
        if (lane == 0 && my_max)
        {
            if (max_count != nullptr)
                atomicMax(max_count, my_max);
            if (sum_count != nullptr)
                atomicAdd(sum_count, my_sum);
        }
    }
}

// upper bound of the minimisers of a read (pair): every window could emit (GC.cpp:690-700); items: its segments (k2_thread.cuh);
// max_windows: the most windows of one mate in the batch
__global__ void k_hash_upper_bounds(const uint32_t *__restrict__ len1, const uint32_t *__restrict__ len2, uint32_t n, uint32_t w, uint32_t *__restrict__ ub,
                                    uint32_t *__restrict__ items, uint32_t *__restrict__ max_windows)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t       u = 0, it = 0, mx = 0;
    if (i < n)
    {
        const uint32_t a = len1[i];
        if (a >= w)
        {
            u  = a - w + 1;
            mx = u;
            it = k2t::segments_of(a, w);
            if (len2 != nullptr && len2[i] >= w)
            {
                u += len2[i] - w + 1;
                mx = max(mx, len2[i] - w + 1);
                it += k2t::segments_of(len2[i], w);
            }
        }
        ub[i] = u;
        if (items != nullptr)
            items[i] = it;
    }
    if (max_windows != nullptr)
    {
        mx = __reduce_max_sync(0xffffffffu, mx);
        if ((threadIdx.x & 31) == 0 && mx)
            atomicMax(max_windows, mx);
    }
}

// ---- K2t over segments: long reads (k2_thread.cuh, segment()) ----------------------------------------------------------------
// One thread per item = (read, mate, segment of kSegWindows windows); item_off[read] = first item of the read.  Hashes go to the
// upper-bound layout (one slot per window) at the slot of the segment's first window, seg_cnt[item] = count | flag.
__global__ void __launch_bounds__(k2t::kThreads)
    k_minimisers_segments(const uint8_t *__restrict__ blk1, const uint32_t *__restrict__ off1, const uint32_t *__restrict__ len1, const uint8_t *__restrict__ blk2,
                          const uint32_t *__restrict__ off2, const uint32_t *__restrict__ len2, uint32_t n_reads, uint32_t k, uint32_t w,
                          const uint64_t *__restrict__ item_off, const uint64_t *__restrict__ hash_off, uint64_t *__restrict__ hashes, uint32_t *__restrict__ seg_cnt)
{
    extern __shared__ __align__(16) uint64_t k2t_smem[];
    k2t::LutEntry *lut = reinterpret_cast<k2t::LutEntry *>(k2t_smem);
    const uint32_t tid = threadIdx.x;
    const uint64_t seed = kMinimiserSeed >> (64 - 2 * k);
    const uint64_t mask = (1ull << (2 * k)) - 1;
    const uint32_t lut_s  = (uint32_t)__cvta_generic_to_shared(k2t_smem);
    const uint32_t ring_s = lut_s + k2t::kLutBytes + tid * 8;
    constexpr uint32_t kStride = k2t::kThreads * 8;
    for (uint32_t c = tid; c < 256; c += k2t::kThreads)
        lut[c] = k2t::lut_entry(c, k);
    __syncthreads();
    const uint64_t n_items = item_off[n_reads];
    for (uint64_t first = (uint64_t)blockIdx.x * k2t::kThreads; first < n_items; first += (uint64_t)gridDim.x * k2t::kThreads)
    {
        const uint64_t item = first + tid;
        if (item >= n_items)
            continue;
        uint32_t lo = 0, hi = n_reads; // the last read whose first item is <= item (reads without items share their successor's)
        while (hi - lo > 1)
        {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (item_off[mid] <= item)
                lo = mid;
            else
                hi = mid;
        }
        const uint32_t read = lo, local = (uint32_t)(item - item_off[read]);
        const uint32_t L1 = len1[read], s1 = k2t::segments_of(L1, w);
        const bool     second = local >= s1;
        const uint8_t *p   = second ? blk2 + off2[read] : blk1 + off1[read];
        const uint32_t L   = second ? len2[read] : L1;
        uint64_t      *out = hashes + hash_off[read] + (second ? L1 - w + 1 : 0u);
        seg_cnt[item]      = k2t::segment(p, L, second ? local - s1 : local, k, w, seed, mask, lut_s, out, ring_s, kStride);
    }
}

// One warp per read: the segments' hashes move left to follow each other from hash_off[read] (in place: a destination is never
// right of its source, and values are read before the lanes store), counts[read] = their sum, flags[read] = a segment was flagged.
__global__ void __launch_bounds__(256)
    k_segments_compact(const uint32_t *__restrict__ len1, uint32_t n_reads, uint32_t w, const uint64_t *__restrict__ item_off, const uint64_t *__restrict__ hash_off,
                       uint64_t *hashes, const uint32_t *__restrict__ seg_cnt, uint32_t *__restrict__ counts, uint8_t *__restrict__ flags)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t read = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; read < n_reads; read += warps)
    {
        const uint64_t i0 = item_off[read], i1 = item_off[read + 1];
        uint32_t       dst = 0, flag = 0;
        if (i1 - i0 == 1)
        { // a short read: one segment, already in place
            const uint32_t c = seg_cnt[i0];
            dst  = c & ~k2t::kSegFlag;
            flag = c >> 31;
        }
        else if (i1 > i0)
        {
            uint64_t *h = hashes + hash_off[read];
            const uint32_t L1 = len1[read], s1 = k2t::segments_of(L1, w), win1 = L1 - w + 1;
            for (uint64_t base = i0; base < i1; base += 32)
            { // 32 segments at a time: counts by the lanes, destinations by a scan
                const uint32_t m   = (uint32_t)min((uint64_t)32, i1 - base);
                const uint32_t raw = lane < m ? seg_cnt[base + lane] : 0u;
                const uint32_t c   = raw & ~k2t::kSegFlag;
                flag |= __ballot_sync(0xffffffffu, raw >> 31) != 0;
                uint32_t incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1)
                {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                    if ((int)lane >= d)
                        incl += t;
                }
                for (uint32_t g = 0; g < m; ++g)
                {
                    const uint32_t cg    = __shfl_sync(0xffffffffu, c, g);
                    const uint32_t dg    = dst + __shfl_sync(0xffffffffu, incl, g) - cg;
                    const uint32_t local = (uint32_t)(base - i0) + g;
                    const uint32_t src   = local < s1 ? local * k2t::kSegWindows : win1 + (local - s1) * k2t::kSegWindows;
                    if (src != dg)
                        for (uint32_t i = 0; i < cg; i += 128)
                        {
                            uint64_t v[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                v[u] = i + u * 32 + lane < cg ? h[src + i + u * 32 + lane] : 0;
                            __syncwarp();
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (i + u * 32 + lane < cg)
                                    h[dg + i + u * 32 + lane] = v[u];
                            __syncwarp();
                        }
                }
                dst += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        if (lane == 0)
        {
            counts[read] = dst;
            flags[read]  = (uint8_t)flag;
        }
    }
}

} // namespace

static int k2_forced();

void launch_minimisers(const uint8_t *blk1, const uint32_t *off1, const uint32_t *len1, const uint8_t *blk2, const uint32_t *off2,
                       const uint32_t *len2, uint32_t n_reads, uint32_t k, uint32_t w, int mode, uint32_t *counts,
                       const uint64_t *hash_off, uint64_t *hashes, uint32_t *max_count, unsigned long long *sum_count, cudaStream_t st)
{
    if (n_reads == 0)
        return;
    const uint32_t W      = w - k + 1;
    // K2t is the default where its parameter range allows (measured on B200, c2: 0.99 ms vs 3.07 ms per 2^21 reads);
    // GANON_B200_K2=warp keeps every read on the warp-per-read kernel, =thread every read K2t can take on K2t (read once per process)
    if (k2_forced() != 1 && k <= k2t::kMaxK && W <= k2t::kMaxW)
    { // K2t: one thread per read, CTAs of 128 reads (k2_thread.cuh)
        int dev_t = 0, sms_t = 148;
        cudaGetDevice(&dev_t);
        cudaDeviceGetAttribute(&sms_t, cudaDevAttrMultiProcessorCount, dev_t);
        const size_t   smem_t = k2t::kLutBytes + (size_t)W * k2t::kThreads * 8; // <= 34 KiB
        const uint32_t want_t = (n_reads + k2t::kThreads - 1) / k2t::kThreads;
        const uint32_t cap_t  = (uint32_t)sms_t * 128;         // short-lived CTAs: the tail of the grid stays small
        const uint32_t grid_t = want_t < cap_t ? want_t : cap_t;
        if (mode == 0)
            k_minimisers_thread<0><<<grid_t, k2t::kThreads, smem_t, st>>>(blk1, off1, len1, blk2, off2, len2, n_reads, k, w, counts, hash_off, hashes, max_count, sum_count, nullptr);
        else if (mode == 1)
            k_minimisers_thread<1><<<grid_t, k2t::kThreads, smem_t, st>>>(blk1, off1, len1, blk2, off2, len2, n_reads, k, w, counts, hash_off, hashes, max_count, sum_count, nullptr);
        else
            k_minimisers_thread<2><<<grid_t, k2t::kThreads, smem_t, st>>>(blk1, off1, len1, blk2, off2, len2, n_reads, k, w, counts, hash_off, hashes, max_count, sum_count, nullptr);
        return;
    }
    const uint32_t nv_cap = (K2_TILE + W + 2 + 1) & ~1u;          // values per warp and level
    const uint32_t nb_cap = (nv_cap + k + 15) & ~15u;             // bases per warp (multiple of 16: keeps 8-byte alignment)
    uint32_t       n_lev  = 0;
    while ((1u << n_lev) <= W)
        ++n_lev;                                                   // levels 0..floor(log2 W)
    if (k <= 26 && W >= 4)
        n_lev = 1;                                                 // run-per-lane path: values only
    const uint32_t sp_cap = (nb_cap >> 5) + 2;
    const size_t   per_w  = ((size_t)n_lev * nv_cap + sp_cap) * 8 + nb_cap;
    uint32_t       wpc    = K2_WARPS;
    while (wpc > 1 && wpc * per_w > 200 * 1024)
        wpc >>= 1;
    const size_t   smem   = (size_t)wpc * per_w;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t want = (n_reads + wpc - 1) / wpc;
    uint32_t       full = (uint32_t)sms * 8; // 64 warps per SM
    const uint32_t fine = (want + 63) / 64;  // ~64 reads per warp: short-lived CTAs (see launch_k3)
    if (fine > full)
        full = fine;
    const uint32_t grid = want < full ? want : full;
#define GNB_K2(M)                                                                                                                     \
    {                                                                                                                                 \
        cudaFuncSetAttribute(k_minimisers<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                                \
        k_minimisers<M><<<grid, wpc * 32, smem, st>>>(blk1, off1, len1, blk2, off2, len2, n_reads, k, w, nv_cap, nb_cap, counts, \
                                                            hash_off, hashes, max_count, sum_count);                                  \
    }
    if (mode == 0)
        GNB_K2(0)
    else if (mode == 1)
        GNB_K2(1)
    else
        GNB_K2(2)
#undef GNB_K2
}

void launch_hash_upper_bounds(const uint32_t *len1, const uint32_t *len2, uint32_t n, uint32_t w, uint32_t *ub, cudaStream_t st, uint32_t *items,
                              uint32_t *max_windows)
{
    if (n)
        k_hash_upper_bounds<<<(n + 255) / 256, 256, 0, st>>>(len1, len2, n, w, ub, items, max_windows);
}

static int k2_forced() // GANON_B200_K2=warp / =thread: one kernel for every read (read once per process); 0: the library picks
{
    static const int forced = [] { const char *e = getenv("GANON_B200_K2"); return e && e[0] == 'w' ? 1 : e && e[0] == 't' ? 2 : 0; }();
    return forced;
}

bool minimisers_segmented(uint32_t k, uint32_t w, uint32_t n_reads, uint32_t max_windows)
{
    if (k2_forced() != 0 || k > k2t::kMaxK || w - k + 1 > k2t::kMaxW)
        return false;
    // a batch of many reads of a few thousand bases fills the GPU with a thread per read; one very long read in it would be its tail
    return max_windows > 2 * k2t::kSegWindows && (n_reads < 131072 || max_windows > 8 * k2t::kSegWindows);
}

uint64_t minimiser_segments_bound(uint64_t total_windows, uint32_t n_reads) { return total_windows / k2t::kSegWindows + 2ull * n_reads; }

void launch_minimisers_segmented(const uint8_t *blk1, const uint32_t *off1, const uint32_t *len1, const uint8_t *blk2, const uint32_t *off2, const uint32_t *len2,
                                 uint32_t n_reads, uint32_t k, uint32_t w, const uint64_t *item_off, uint64_t item_bound, uint32_t *seg_cnt, uint8_t *flags,
                                 uint32_t *counts, const uint64_t *hash_off, uint64_t *hashes, uint32_t *max_count, unsigned long long *sum_count, cudaStream_t st)
{
    if (n_reads == 0)
        return;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t W    = w - k + 1;
    const size_t   smem = k2t::kLutBytes + (size_t)W * k2t::kThreads * 8;
    const uint64_t want = (item_bound + k2t::kThreads - 1) / k2t::kThreads;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)sms * 128);
    k_minimisers_segments<<<grid, k2t::kThreads, smem, st>>>(blk1, off1, len1, blk2, off2, len2, n_reads, k, w, item_off, hash_off, hashes, seg_cnt);
    const uint32_t grid_c = (uint32_t)std::min<uint64_t>(((uint64_t)n_reads * 32 + 255) / 256, (uint64_t)sms * 64);
    k_segments_compact<<<grid_c, 256, 0, st>>>(len1, n_reads, w, item_off, hash_off, hashes, seg_cnt, counts, flags);
    // reads with a flagged segment (repeats all along a warm-up) are walked from the start by one thread; the pass also sums up
    const uint32_t want_t = (n_reads + k2t::kThreads - 1) / k2t::kThreads;
    const uint32_t grid_t = std::min<uint32_t>(want_t, (uint32_t)sms * 128);
    k_minimisers_thread<2><<<grid_t, k2t::kThreads, smem, st>>>(blk1, off1, len1, blk2, off2, len2, n_reads, k, w, counts, hash_off, hashes, max_count, sum_count, flags);
}

// ---------------------------------------------------------------------------------------------------------------------
// counts -> offsets.  Reads with more than 65535 minimisers (GC.cpp:674,706) keep their slot; K3 skips them.
// ---------------------------------------------------------------------------------------------------------------------
namespace
{
__global__ void k_widen_counts(const uint32_t *__restrict__ counts, uint64_t *__restrict__ wide, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        wide[i] = counts[i];
    if (i == n)
        wide[i] = 0;
}
} // namespace

size_t scan_tmp_bytes(uint32_t n_reads)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (uint64_t *)nullptr, (uint64_t *)nullptr, (int)(n_reads + 1));
    return bytes + 256;
}

void launch_scan_counts(const uint32_t *counts, uint64_t *hash_off, uint32_t n_reads, void *tmp, size_t tmp_bytes, cudaStream_t st)
{
    // widen in place into hash_off, then scan hash_off -> hash_off (cub allows in-place exclusive scans)
    k_widen_counts<<<(n_reads + 1 + 255) / 256, 256, 0, st>>>(counts, hash_off, n_reads);
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, hash_off, hash_off, (int)(n_reads + 1), st);
}

// =====================================================================================================================
// K3 -- IBF count
// =====================================================================================================================
namespace
{

constexpr int K3_WARPS = 8;

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (a & c) | (b & c); } // one LOP3

// 16-byte streaming row load: read-only path, do not keep the line in L1 (rows are touched once per lookup)
__device__ __forceinline__ uint4 ldg_stream16(const uint64_t *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// plain 8-byte load as a volatile asm: keeps the issue order written in the source (used to software-pipeline the
// minimiser fetch ahead of the row gathers)
__device__ __forceinline__ uint64_t ldg_u64(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg_stream8(const uint64_t *p)
{
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// Add four 1-bit-per-bin masks to the bit-sliced counters P (plane j = bit j of every bin's count):
// two full adders at weight 1, one at weight 2, then a ripple of the weight-4 carry.
template <int NP>
__device__ __forceinline__ void csa_add4(uint32_t (&P)[NP][4], const uint32_t (&x)[4][4])
{
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const uint32_t a = x[0][r], b = x[1][r], c = x[2][r], d = x[3][r];
        const uint32_t p0 = P[0][r];
        const uint32_t s1 = p0 ^ a ^ b, c1 = maj3(p0, a, b);
        const uint32_t s2 = s1 ^ c ^ d, c2 = maj3(s1, c, d);
        P[0][r]           = s2;
        const uint32_t p1 = P[1][r];
        P[1][r]           = p1 ^ c1 ^ c2;
        uint32_t carry    = maj3(p1, c1, c2);
#pragma unroll
        for (int j = 2; j < NP; ++j)
        {
            const uint32_t t = P[j][r] & carry;
            P[j][r] ^= carry;
            carry = t;
        }
    }
}

template <int NP>
__device__ __forceinline__ uint32_t sliced_ge(const uint32_t (&P)[NP][4], int r, uint32_t T)
{
    uint32_t ge = 0xffffffffu; // "equal so far" counts as >=
#pragma unroll
    for (int j = 0; j < NP; ++j)
        ge = ((T >> j) & 1u) ? (P[j][r] & ge) : (P[j][r] | ge);
    return ge;
}

template <int NP>
__device__ __forceinline__ uint32_t sliced_get(const uint32_t (&P)[NP][4], int r, uint32_t bit)
{
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < NP; ++j)
        c |= ((P[j][r] >> bit) & 1u) << j;
    return c;
}

template <int NP>
__device__ __forceinline__ uint32_t sliced_sum(const uint32_t (&P)[NP][4], uint32_t reg, uint32_t mask)
{
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < NP; ++j)
    {
        const uint32_t pj = reg == 0 ? P[j][0] : reg == 1 ? P[j][1] : reg == 2 ? P[j][2] : P[j][3];
        s += (uint32_t)__popc(pj & mask) << j;
    }
    return s;
}

// MODE 0: sparse tuples; MODE 1: dense counts (test hook); MODE 2: HIBF item (merged bins feed the next worklist)
constexpr uint32_t kMergedBin = 0x80000000u; // bin_node flag: the bin is a merged bin, low bits = child IBF

struct WorkOut
{
    uint64_t           *tuples;
    unsigned long long *cursor;
    uint64_t            cap;
    uint16_t           *dense;
    uint2              *items; // MODE 2: (read, child ibf) for merged bins that reach the threshold
    unsigned long long *items_cursor;
    uint64_t            items_cap;
    unsigned long long *bytes; // MODE 2: algorithmic bytes of the round (sum over items of n_hashes * h * row_words * 8)
};

// One (read, chunk) work item: gather + AND + bit-sliced count over the read's minimisers, then the epilogue.
template <int H, int NP, bool ALIGNED, int MODE>
__device__ __forceinline__ void count_item(const IbfDev &f, uint32_t read, uint32_t chunk, const uint64_t *__restrict__ hashes,
                                           const uint64_t *__restrict__ hash_off, const uint32_t *__restrict__ counts, double rel_cutoff,
                                           const WorkOut &wo, uint64_t *s_row, uint32_t lane)
{
    constexpr int  PER    = 32 / H; // minimisers whose rows are staged per round
    const uint64_t seed_l = ibf_seed(lane % H);
    uint64_t *const tuples = wo.tuples;
    unsigned long long *const cursor = wo.cursor;
    const uint64_t  cap   = wo.cap;
    uint16_t *const dense = wo.dense;
    const uint64_t h0 = hash_off[read];
    const uint64_t nn = counts != nullptr ? (uint64_t)counts[read] : hash_off[read + 1] - h0; // upper-bound vs exact layout
    if (nn == 0 || nn > 65535) // skipped: shorter than the window / more minimisers than the counter type holds
        return;
    const uint32_t n  = (uint32_t)nn;
    const uint32_t w0 = chunk * 64 + lane * 2; // first of the lane's two bin-words inside the row
    const bool     v0 = w0 < f.row_words, v1 = (w0 + 1) < f.row_words;
    const uint64_t *lane_base = f.data + w0;

    uint32_t P[NP][4];
#pragma unroll
    for (int j = 0; j < NP; ++j)
        P[j][0] = P[j][1] = P[j][2] = P[j][3] = 0;

    for (uint32_t base = 0; base < n; base += PER)
    {
        __syncwarp();
        if (lane < PER * H)
        {
            const uint32_t m = base + lane / H;
            if (m < n)
                s_row[lane] = ibf_row(hashes[h0 + m], seed_l, f.hash_shift, f.bin_size) * f.row_words;
        }
        __syncwarp();
        const uint32_t cnt = min((uint32_t)PER, n - base);
        for (uint32_t jb = 0; jb < cnt; jb += 4)
        {
            uint32_t x[4][4];
            uint4    rows[4][H];
            // issue every load of the block first (up to 4*H 128-bit loads in flight per lane)
#pragma unroll
            for (int q = 0; q < 4; ++q)
            {
                const bool ok = (jb + q) < cnt;
#pragma unroll
                for (int i = 0; i < H; ++i)
                {
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (ok)
                    {
                        const uint64_t *p = lane_base + s_row[(jb + q) * H + i];
                        if (ALIGNED)
                        {
                            if (v0)
                                v = ldg_stream16(p);
                        }
                        else
                        {
                            if (v0)
                            {
                                const uint2 a = ldg_stream8(p);
                                v.x = a.x;
                                v.y = a.y;
                            }
                            if (v1)
                            {
                                const uint2 b = ldg_stream8(p + 1);
                                v.z = b.x;
                                v.w = b.y;
                            }
                        }
                    }
                    rows[q][i] = v;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
            {
                uint4 a = rows[q][0];
#pragma unroll
                for (int i = 1; i < H; ++i)
                {
                    a.x &= rows[q][i].x;
                    a.y &= rows[q][i].y;
                    a.z &= rows[q][i].z;
                    a.w &= rows[q][i].w;
                }
                x[q][0] = a.x;
                x[q][1] = a.y;
                x[q][2] = a.z;
                x[q][3] = a.w;
            }
            csa_add4<NP>(P, x);
        }
    }

    if (MODE == 1)
    {
        // dense dump: counts[read][bin]
        uint16_t *o = dense + (uint64_t)read * ((uint64_t)f.row_words * 64) + (uint64_t)w0 * 64;
#pragma unroll
        for (int r = 0; r < 4; ++r)
        {
            const bool valid = (r < 2) ? v0 : v1;
            if (valid)
                for (uint32_t b = 0; b < 32; ++b)
                    o[r * 32 + b] = (uint16_t)sliced_get<NP>(P, r, b);
        }
        return;
    }

    // ---- select_matches (GC.cpp:504-541) on the lane's 128 bins ----
    const uint32_t T = threshold_cutoff(n, rel_cutoff);
    uint32_t       cand[4];
    uint32_t       mine = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
        const bool valid = (r < 2) ? v0 : v1;
        cand[r] = valid ? (sliced_ge<NP>(P, r, T) & f.single_mask[(uint32_t)w0 * 2 + r]) : 0u;
        mine += __popc(cand[r]);
    }
    if (__any_sync(0xffffffffu, mine != 0))
    {
        // warp-aggregated append: one atomic per warp
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d)
                incl += y;
        }
        const uint32_t     total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned long long start = 0;
        if (MODE != 2 && lane == 0)
            start = atomicAdd(cursor, (unsigned long long)total);
        start = __shfl_sync(0xffffffffu, start, 0);
        uint64_t pos = start + (incl - mine);
#pragma unroll
        for (int r = 0; r < 4; ++r)
        {
            uint32_t m = cand[r];
            while (m)
            {
                const uint32_t b = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t node = f.bin_node[(uint32_t)w0 * 64 + r * 32 + b];
                if (MODE == 2)
                {
                    if (node & kMergedBin)
                    { // descend: the child IBF is queried with the full hash list (HIBF.hpp:447-451)
                        const unsigned long long ip = atomicAdd(wo.items_cursor, 1ULL);
                        if (ip < wo.items_cap)
                            wo.items[ip] = make_uint2(read, node & ~kMergedBin);
                    }
                    else
                    {
                        const unsigned long long tp = atomicAdd(cursor, 1ULL);
                        if (tp < cap)
                            tuples[tp] = make_tuple64(read, node, 0, sliced_get<NP>(P, r, b));
                    }
                    continue;
                }
                if (pos < cap)
                    tuples[pos] = make_tuple64(read, node, 0, sliced_get<NP>(P, r, b));
                ++pos;
            }
        }
    }
    // ---- nodes made of several bins: masked popcount sums per segment ----
    if (f.seg_off != nullptr)
    {
        const uint32_t s0 = f.seg_off[chunk * 32 + lane], s1 = f.seg_off[chunk * 32 + lane + 1];
        for (uint32_t s = s0; s < s1; ++s)
        {
            const Seg      sg  = f.segs[s];
            uint32_t       sum = sliced_sum<NP>(P, sg.reg, sg.mask);
            uint32_t       partial = 1;
            if (sg.complete == 2)
            { // HIBF user bin split inside this register: u16 running sum, then the threshold (HIBF.hpp:437-458)
                sum &= 0xFFFF;
                partial = 0;
                if (sum < T)
                    sum = 0;
            }
            else if (sg.complete)
            {
                sum     = min(sum, n); // GC.cpp:525-526
                partial = 0;
                if (sum < T)
                    sum = 0;
            }
            if (sum != 0)
            {
                const unsigned long long pos = atomicAdd(cursor, 1ULL);
                if (pos < cap)
                    tuples[pos] = make_tuple64(read, sg.node, partial, sum);
            }
        }
    }

}

template <int H, int NP, bool ALIGNED, int MODE>
__global__ void __launch_bounds__(K3_WARPS * 32)
    k_ibf_count(const IbfDev f, const uint64_t *__restrict__ hashes, const uint64_t *__restrict__ hash_off, const uint32_t *__restrict__ counts,
                const uint8_t *__restrict__ active, uint32_t n_reads, double rel_cutoff, const WorkOut wo)
{
    __shared__ uint64_t s_row[K3_WARPS][32];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t n_items = (uint64_t)n_reads * f.n_chunks;
    const uint64_t stride  = (uint64_t)gridDim.x * K3_WARPS;
    for (uint64_t item = (uint64_t)blockIdx.x * K3_WARPS + wib; item < n_items; item += stride)
    {
        const uint32_t read  = (uint32_t)(item / f.n_chunks);
        const uint32_t chunk = (uint32_t)(item - (uint64_t)read * f.n_chunks);
        if (active != nullptr && active[read] == 0)
            continue;
        count_item<H, NP, ALIGNED, MODE>(f, read, chunk, hashes, hash_off, counts, rel_cutoff, wo, s_row[wib], lane);
    }
}

// HIBF traversal round (bulk_count_impl, HIBF.hpp:433-460, level-synchronous): one warp per (read, ibf) item, all
// chunks of that sub-IBF in turn.  User bins reaching the threshold become tuples, merged bins become items of the
// next round.
template <int H, int NP>
__global__ void __launch_bounds__(K3_WARPS * 32)
    k_hibf_count(const IbfDev *__restrict__ table, const uint2 *__restrict__ items, uint32_t n_items, const uint64_t *__restrict__ hashes,
                 const uint64_t *__restrict__ hash_off, const uint32_t *__restrict__ counts, double rel_cutoff, const WorkOut wo)
{
    __shared__ uint64_t s_row[K3_WARPS][32];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned long long bytes = 0;
    for (uint32_t it = blockIdx.x * K3_WARPS + wib; it < n_items; it += gridDim.x * K3_WARPS)
    {
        const uint2  item = items[it];
        const IbfDev f    = table[item.y];
        bytes += (unsigned long long)counts[item.x] * f.hash_funs * f.row_words * 8;
        for (uint32_t chunk = 0; chunk < f.n_chunks; ++chunk)
            count_item<H, NP, false, 2>(f, item.x, chunk, hashes, hash_off, counts, rel_cutoff, wo, s_row[wib], lane);
    }
    if (lane == 0 && bytes && wo.bytes)
        atomicAdd(wo.bytes, bytes);
}

// Sub-IBFs of an HIBF are narrow (raptor's t_max: 64 .. ~2048 technical bins = 1 .. 32 words per row), so a whole warp
// per (read, ibf) item leaves most lanes idle and -- worse -- leaves one short dependent-load chain per warp.  Here an
// item takes G lanes (lane `sub` owns words 2*sub, 2*sub+1 of the row, as in count_item) and a warp works on 32/G items
// at once: every lane hashes its item's minimisers itself, gathers its 16 bytes of the h rows, ANDs and counts in its
// own bit-sliced planes; no cross-lane traffic until the warp-aggregated append of the results.
template <int H, int NP, int G>
__global__ void __launch_bounds__(K3_WARPS * 32, 2)
    k_hibf_count_narrow(const IbfDev *__restrict__ table, const uint2 *__restrict__ items, uint32_t n_items, const uint64_t *__restrict__ hashes,
                        const uint64_t *__restrict__ hash_off, const uint32_t *__restrict__ counts, double rel_cutoff, const WorkOut wo)
{
    constexpr uint32_t IPW  = 32 / G; // items per warp
    const uint32_t     lane = threadIdx.x & 31, lt = (1u << lane) - 1, sub = lane % G;
    const uint32_t     warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long bytes = 0;
    for (uint32_t base = warp * IPW; base < n_items; base += n_warps * IPW)
    {
        const uint32_t it   = base + lane / G;
        const bool     have = it < n_items;
        const uint2    item = have ? items[it] : make_uint2(0, 0);
        const IbfDev  *fp   = table + item.y;
        // every field up front: the loads overlap with the gather loop instead of forming a chain in the epilogue
        const uint32_t row_words = fp->row_words, hash_shift = fp->hash_shift;
        const uint64_t bin_size = fp->bin_size;
        const uint32_t *const f_single = fp->single_mask, *const f_bin_node = fp->bin_node, *const f_seg_off = fp->seg_off;
        const Seg *const      f_segs = fp->segs;
        uint32_t       n  = 0;
        uint64_t       h0 = 0;
        if (have)
        {
            const uint32_t nn = counts[item.x];
            if (nn > 0 && nn <= 65535)
            {
                n  = nn;
                h0 = hash_off[item.x];
            }
        }
        const uint32_t w0 = sub * 2;
        const bool     v0 = n != 0 && w0 < row_words, v1 = n != 0 && (w0 + 1) < row_words;
        const bool     al = (row_words & 1) == 0; // 16-byte loads stay aligned
        const uint64_t *lane_base = fp->data + w0;
        if (sub == 0)
            bytes += (unsigned long long)n * H * row_words * 8;

        uint32_t P[NP][4];
#pragma unroll
        for (int j = 0; j < NP; ++j)
            P[j][0] = P[j][1] = P[j][2] = P[j][3] = 0;
        // With G > 1 the G lanes of an item would each hash the same 4 x H (minimiser, hash function) pairs of a block: 64-bit
        // multiplies that made up half of the kernel's instructions (ncu, round 2: 1050 warp-instructions per block, issue
        // slots 49 % busy at 4 warps per scheduler).  Instead every pair is hashed by ONE lane of the group and handed round
        // with shuffles inside the group (all lanes of a group share the item, hence the trip count of the loop below; the
        // loop is entered by the whole group and only the loads depend on the lane's columns).
        constexpr int      PAIRS = 4 * H, PER = (PAIRS + G - 1) / G;
        const uint32_t     gmask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (lane & ~(uint32_t)(G - 1)));
        if (G > 1 ? n != 0 : v0)
        {
            // the four minimisers of the next block are fetched while the rows of the current one are in flight
            uint64_t xn[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                xn[q] = (uint32_t)q < n ? ldg_u64(hashes + h0 + q) : 0;
            for (uint32_t m = 0; m < n; m += 4)
            {
                uint4    rows[4][H];
                uint64_t xc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    xc[q] = xn[q];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    xn[q] = (m + 4 + q) < n ? ldg_u64(hashes + h0 + m + 4 + q) : 0;
                uint64_t mine[PER]; // word offsets of the rows this lane hashed for its group
                if (G > 1)
                {
#pragma unroll
                    for (int j = 0; j < PER; ++j)
                    {
                        const uint32_t pr = sub + (uint32_t)j * G, q = pr / H, i = pr % H;
                        const uint64_t x  = q == 0 ? xc[0] : q == 1 ? xc[1] : q == 2 ? xc[2] : xc[3];
                        mine[j]           = pr < (uint32_t)PAIRS ? ibf_row(x, ibf_seed(i), hash_shift, bin_size) * row_words : 0;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                {
                    const bool     ok = (m + q) < n;
                    const uint64_t x  = xc[q];
#pragma unroll
                    for (int i = 0; i < H; ++i)
                    {
                        uint64_t off;
                        if (G > 1)
                        {
                            constexpr int dummy = 0;
                            (void)dummy;
                            const int      pr = q * H + i;
                            const uint64_t mv = mine[pr / G];
                            const uint32_t lo = __shfl_sync(gmask, (uint32_t)mv, pr % G, G), hi = __shfl_sync(gmask, (uint32_t)(mv >> 32), pr % G, G);
                            off               = ((uint64_t)hi << 32) | lo;
                        }
                        else
                            off = ibf_row(x, ibf_seed(i), hash_shift, bin_size) * row_words;
                        uint4 v = make_uint4(0, 0, 0, 0);
                        if (ok && v0)
                        {
                            const uint64_t *p = lane_base + off;
                            if (al)
                                v = ldg_stream16(p);
                            else
                            {
                                const uint2 a = ldg_stream8(p);
                                v.x = a.x;
                                v.y = a.y;
                                if (v1)
                                {
                                    const uint2 b = ldg_stream8(p + 1);
                                    v.z = b.x;
                                    v.w = b.y;
                                }
                            }
                        }
                        rows[q][i] = v;
                    }
                }
                uint32_t x4[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                {
                    uint4 a = rows[q][0];
#pragma unroll
                    for (int i = 1; i < H; ++i)
                    {
                        a.x &= rows[q][i].x;
                        a.y &= rows[q][i].y;
                        a.z &= rows[q][i].z;
                        a.w &= rows[q][i].w;
                    }
                    x4[q][0] = a.x;
                    x4[q][1] = a.y;
                    x4[q][2] = a.z;
                    x4[q][3] = a.w;
                }
                csa_add4<NP>(P, x4);
            }
        }

        // ---- epilogue: bins reaching the threshold -> tuples (user bins) / items of the next round (merged bins) ----
        const uint32_t T = threshold_cutoff(n, rel_cutoff);
        uint32_t       cand[4];
        uint32_t       nt = 0, ni = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r)
        {
            const bool valid = (r < 2) ? v0 : v1;
            cand[r] = valid ? (sliced_ge<NP>(P, r, T) & f_single[w0 * 2 + r]) : 0u;
            uint32_t mm = cand[r];
            while (mm)
            {
                const uint32_t b = __ffs(mm) - 1;
                mm &= mm - 1;
                if (f_bin_node[w0 * 64 + r * 32 + b] & kMergedBin)
                    ++ni;
                else
                    ++nt;
            }
        }
        uint32_t s0 = 0, s1 = 0;
        if (v0 && f_seg_off != nullptr)
        {
            s0 = f_seg_off[sub];
            s1 = f_seg_off[sub + 1];
            for (uint32_t sI = s0; sI < s1; ++sI)
            {
                const Seg sg = f_segs[sI];
                uint32_t  sum = sliced_sum<NP>(P, sg.reg, sg.mask);
                if (sg.complete == 2)
                    sum = (sum & 0xFFFF) < T ? 0 : (sum & 0xFFFF);
                else if (sg.complete)
                    sum = min(sum, n) < T ? 0 : min(sum, n);
                nt += sum != 0;
            }
        }
        if (!__any_sync(0xffffffffu, (nt | ni) != 0))
            continue;
        uint32_t it_incl = nt, ii_incl = ni;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t a = __shfl_up_sync(0xffffffffu, it_incl, d), b = __shfl_up_sync(0xffffffffu, ii_incl, d);
            if (lane >= (uint32_t)d)
            {
                it_incl += a;
                ii_incl += b;
            }
        }
        const uint32_t     tot_t = __shfl_sync(0xffffffffu, it_incl, 31), tot_i = __shfl_sync(0xffffffffu, ii_incl, 31);
        unsigned long long bt = 0, bi = 0;
        if (lane == 0)
        {
            if (tot_t)
                bt = atomicAdd(wo.cursor, (unsigned long long)tot_t);
            if (tot_i)
                bi = atomicAdd(wo.items_cursor, (unsigned long long)tot_i);
        }
        bt = __shfl_sync(0xffffffffu, bt, 0) + (it_incl - nt);
        bi = __shfl_sync(0xffffffffu, bi, 0) + (ii_incl - ni);
        (void)lt;
#pragma unroll
        for (int r = 0; r < 4; ++r)
        {
            uint32_t mm = cand[r];
            while (mm)
            {
                const uint32_t b = __ffs(mm) - 1;
                mm &= mm - 1;
                const uint32_t node = f_bin_node[w0 * 64 + r * 32 + b];
                if (node & kMergedBin)
                {
                    if (bi < wo.items_cap)
                        wo.items[bi] = make_uint2(item.x, node & ~kMergedBin);
                    ++bi;
                }
                else
                {
                    if (bt < wo.cap)
                        wo.tuples[bt] = make_tuple64(item.x, node, 0, sliced_get<NP>(P, r, b));
                    ++bt;
                }
            }
        }
        for (uint32_t sI = s0; sI < s1; ++sI)
        {
            const Seg sg = f_segs[sI];
            uint32_t  sum = sliced_sum<NP>(P, sg.reg, sg.mask), partial = 1;
            if (sg.complete == 2)
            {
                sum     = (sum & 0xFFFF) < T ? 0 : (sum & 0xFFFF);
                partial = 0;
            }
            else if (sg.complete)
            {
                sum     = min(sum, n) < T ? 0 : min(sum, n);
                partial = 0;
            }
            if (sum != 0)
            {
                if (bt < wo.cap)
                    wo.tuples[bt] = make_tuple64(item.x, sg.node, partial, sum);
                ++bt;
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        bytes += __shfl_xor_sync(0xffffffffu, bytes, d);
    if (lane == 0 && bytes && wo.bytes)
        atomicAdd(wo.bytes, bytes);
}

// first round of the traversal: one item (read, top-level IBF) per active read with 1..65535 minimisers
__global__ void k_hibf_seed_items(const uint8_t *__restrict__ active, const uint32_t *__restrict__ counts, uint32_t n_reads, uint2 *__restrict__ items,
                                  unsigned long long *__restrict__ cursor)
{
    const uint32_t r    = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const bool     take = r < n_reads && (active == nullptr || active[r] != 0) && counts[r] > 0 && counts[r] <= 65535;
    const uint32_t m    = __ballot_sync(0xffffffffu, take);
    if (!m)
        return;
    unsigned long long base = 0;
    if (lane == 0)
        base = atomicAdd(cursor, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (take)
        items[base + __popc(m & ((1u << lane) - 1))] = make_uint2(r, 0);
}

template <int H, int NP, bool ALIGNED, int MODE>
void launch_k3(const IbfDev &f, const uint64_t *hashes, const uint64_t *hash_off, const uint32_t *counts, const uint8_t *active, uint32_t n_reads,
               double rel_cutoff, uint64_t *tuples, unsigned long long *cursor, uint64_t cap, uint16_t *dense, cudaStream_t st)
{
    auto     kern  = k_ibf_count<H, NP, ALIGNED, MODE>;
    int      occ   = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, K3_WARPS * 32, 0);
    if (occ < 1)
        occ = 1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t items = (uint64_t)n_reads * f.n_chunks;
    uint64_t       want  = (items + K3_WARPS - 1) / K3_WARPS;
    // at least one resident wave; beyond that CTAs of ~16 items per warp, so that a CTA lives ~0.2 ms and the
    // high-priority staging stream of the next batch (H2D + K1) gets SMs without waiting for the whole launch
    uint64_t       full  = (uint64_t)sms * occ;
    const uint64_t fine  = (want + 15) / 16;
    if (fine > full)
        full = fine;
    const uint32_t grid  = (uint32_t)(want < full ? want : full);
    WorkOut wo{tuples, cursor, cap, dense, nullptr, nullptr, 0, nullptr};
    kern<<<grid, K3_WARPS * 32, 0, st>>>(f, hashes, hash_off, counts, active, n_reads, rel_cutoff, wo);
}

template <int H, int MODE>
void dispatch_k3(const IbfDev &f, const uint64_t *hashes, const uint64_t *hash_off, const uint32_t *counts, const uint8_t *active, uint32_t n_reads,
                 uint32_t max_hashes, double rel_cutoff, uint64_t *tuples, unsigned long long *cursor, uint64_t cap,
                 uint16_t *dense, cudaStream_t st)
{
    const bool aligned = (f.row_words % 2 == 0) && (((uintptr_t)f.data & 15) == 0);
    const bool small   = max_hashes < 256; // 8 planes hold counts up to 255
#define GNB_K3(NPV, AL) launch_k3<H, NPV, AL, MODE>(f, hashes, hash_off, counts, active, n_reads, rel_cutoff, tuples, cursor, cap, dense, st)
    if (small)
    {
        if (aligned)
            GNB_K3(8, true);
        else
            GNB_K3(8, false);
    }
    else
    {
        if (aligned)
            GNB_K3(16, true);
        else
            GNB_K3(16, false);
    }
#undef GNB_K3
}

template <int MODE>
void dispatch_k3_h(const IbfDev &f, const uint64_t *hashes, const uint64_t *hash_off, const uint32_t *counts, const uint8_t *active, uint32_t n_reads,
                   uint32_t max_hashes, double rel_cutoff, uint64_t *tuples, unsigned long long *cursor, uint64_t cap,
                   uint16_t *dense, cudaStream_t st)
{
    switch (f.hash_funs)
    {
    case 1: dispatch_k3<1, MODE>(f, hashes, hash_off, counts, active, n_reads, max_hashes, rel_cutoff, tuples, cursor, cap, dense, st); break;
    case 2: dispatch_k3<2, MODE>(f, hashes, hash_off, counts, active, n_reads, max_hashes, rel_cutoff, tuples, cursor, cap, dense, st); break;
    case 3: dispatch_k3<3, MODE>(f, hashes, hash_off, counts, active, n_reads, max_hashes, rel_cutoff, tuples, cursor, cap, dense, st); break;
    case 4: dispatch_k3<4, MODE>(f, hashes, hash_off, counts, active, n_reads, max_hashes, rel_cutoff, tuples, cursor, cap, dense, st); break;
    default: dispatch_k3<5, MODE>(f, hashes, hash_off, counts, active, n_reads, max_hashes, rel_cutoff, tuples, cursor, cap, dense, st); break;
    }
}

} // namespace

void launch_ibf_count(const IbfDev &f, const uint64_t *hashes, const uint64_t *hash_off, const uint32_t *counts, const uint8_t *active, uint32_t n_reads,
                      uint32_t max_hashes, double rel_cutoff, uint64_t *tuples, unsigned long long *cursor, uint64_t cap, cudaStream_t st)
{
    if (n_reads == 0)
        return;
    dispatch_k3_h<0>(f, hashes, hash_off, counts, active, n_reads, max_hashes, rel_cutoff, tuples, cursor, cap, nullptr, st);
}

void launch_hibf_round(const IbfDev *table, uint32_t hash_funs, const uint2 *items, uint32_t n_items, const uint64_t *hashes, const uint64_t *hash_off,
                       const uint32_t *counts, uint32_t max_hashes, double rel_cutoff, uint64_t *tuples, unsigned long long *cursor, uint64_t cap, uint2 *items_out,
                       unsigned long long *items_cursor, uint64_t items_cap, unsigned long long *bytes, uint32_t lanes_per_item, cudaStream_t st)
{
    if (n_items == 0)
        return;
    WorkOut        wo{tuples, cursor, cap, nullptr, items_out, items_cursor, items_cap, bytes};
    const bool     small = max_hashes < 256;
    if (lanes_per_item >= 1 && lanes_per_item <= 16)
    { // narrow sub-IBFs: G lanes per item
        const uint64_t threads = (uint64_t)n_items * lanes_per_item;
        const uint32_t want_n  = (uint32_t)((threads + K3_WARPS * 32 - 1) / (K3_WARPS * 32));
        const uint32_t grid_n  = want_n < 148u * 8u ? want_n : 148u * 8u;
#define GNB_NARROW(HV, GV)                                                                                                                                        \
    if (small)                                                                                                                                                    \
        k_hibf_count_narrow<HV, 8, GV><<<grid_n, K3_WARPS * 32, 0, st>>>(table, items, n_items, hashes, hash_off, counts, rel_cutoff, wo);                    \
    else                                                                                                                                                          \
        k_hibf_count_narrow<HV, 16, GV><<<grid_n, K3_WARPS * 32, 0, st>>>(table, items, n_items, hashes, hash_off, counts, rel_cutoff, wo);
#define GNB_NARROW_G(HV)                  \
    switch (lanes_per_item)               \
    {                                     \
    case 1: GNB_NARROW(HV, 1) break;      \
    case 2: GNB_NARROW(HV, 2) break;      \
    case 4: GNB_NARROW(HV, 4) break;      \
    case 8: GNB_NARROW(HV, 8) break;      \
    default: GNB_NARROW(HV, 16) break;    \
    }
        switch (hash_funs)
        {
        case 1: GNB_NARROW_G(1) break;
        case 2: GNB_NARROW_G(2) break;
        case 3: GNB_NARROW_G(3) break;
        case 4: GNB_NARROW_G(4) break;
        default: GNB_NARROW_G(5) break;
        }
#undef GNB_NARROW_G
#undef GNB_NARROW
        return;
    }
    const uint32_t want = (n_items + K3_WARPS - 1) / K3_WARPS;
    const uint32_t grid = want < 148u * 4u ? want : 148u * 4u;
#define GNB_HIBF(HV)                                                                                                             \
    if (small)                                                                                                                   \
        k_hibf_count<HV, 8><<<grid, K3_WARPS * 32, 0, st>>>(table, items, n_items, hashes, hash_off, counts, rel_cutoff, wo);           \
    else                                                                                                                         \
        k_hibf_count<HV, 16><<<grid, K3_WARPS * 32, 0, st>>>(table, items, n_items, hashes, hash_off, counts, rel_cutoff, wo);
    switch (hash_funs)
    {
    case 1: GNB_HIBF(1) break;
    case 2: GNB_HIBF(2) break;
    case 3: GNB_HIBF(3) break;
    case 4: GNB_HIBF(4) break;
    default: GNB_HIBF(5) break;
    }
#undef GNB_HIBF
}

void launch_hibf_seed_items(const uint8_t *active, const uint32_t *counts, uint32_t n_reads, uint2 *items, unsigned long long *cursor, cudaStream_t st)
{
    if (n_reads == 0)
        return;
    k_hibf_seed_items<<<(n_reads + 255) / 256, 256, 0, st>>>(active, counts, n_reads, items, cursor);
}

void launch_ibf_count_dense(const IbfDev &f, const uint64_t *hashes, const uint64_t *hash_off, uint32_t n_reads, uint32_t max_hashes,
                            uint16_t *counts, cudaStream_t st)
{
    if (n_reads == 0)
        return;
    dispatch_k3_h<1>(f, hashes, hash_off, nullptr, nullptr, n_reads, max_hashes, 0.0, nullptr, nullptr, 0, counts, st);
}

// ---------------------------------------------------------------------------------------------------------------------
// tuple sort by (read, node): bits [17, 64)
// ---------------------------------------------------------------------------------------------------------------------
size_t sort_tmp_bytes(uint64_t n)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int64_t)n, 16, 64);
    return bytes + 256;
}

void launch_sort_tuples(const uint64_t *in, uint64_t *out, uint64_t n, void *tmp, size_t tmp_bytes, cudaStream_t st)
{
    if (n == 0)
        return;
    cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, in, out, (int64_t)n, 16, 64, st);
}

// ---------------------------------------------------------------------------------------------------------------------
// distinct values of a hash list (ganon-build: the minimiser set of a file): radix sort + unique
// ---------------------------------------------------------------------------------------------------------------------
size_t unique_tmp_bytes(uint64_t n)
{
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, a, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int64_t)n, 0, 64);
    cub::DeviceSelect::Unique(nullptr, b, (const uint64_t *)nullptr, (uint64_t *)nullptr, (unsigned long long *)nullptr, (int64_t)n);
    return std::max(a, b) + 256;
}

void launch_sort_unique(const uint64_t *in, uint64_t *tmp_keys, uint64_t *out, uint64_t n, unsigned long long *d_n_out, void *tmp, size_t tmp_bytes, cudaStream_t st)
{
    if (n == 0)
    {
        cudaMemsetAsync(d_n_out, 0, 8, st);
        return;
    }
    size_t bytes = tmp_bytes;
    cub::DeviceRadixSort::SortKeys(tmp, bytes, in, tmp_keys, (int64_t)n, 0, 64, st);
    bytes = tmp_bytes;
    cub::DeviceSelect::Unique(tmp, bytes, tmp_keys, out, d_n_out, (int64_t)n, st);
}

// =====================================================================================================================
// K4: finishing stage of one hierarchy level on the device (levels with one filter).
//   k_tuple_starts    first tuple of every read in the sorted tuple list
//   k_finish_select   thread per read.  Pass 1: runs of equal (read, node) are summed (partial sums of targets spread over
//                     several registers / chunks / technical bins), capped and compared with the cutoff exactly as
//                     select_matches does (GC.cpp:504-541, 543-577; HIBF running sum wraps at 16 bits, HIBF.hpp:437-441);
//                     max / min of the accepted counts.  Pass 2: rel-filter threshold (GC.cpp:757-758, 579-587) and
//                     --fpr-query (GC.cpp:588-601) per accepted match, LCA of the kept targets (GC.cpp:615-627), sizes of
//                     the read's .all/.one/.unc lines, per-batch totals (struct Total GC.cpp:162-177).
//                     No side effect outside the batch's scratch: the host may still discard the pass (see below).
//   k_finish_write    thread per read, after the exclusive scan of the sizes: report counters (struct Rep GC.cpp:153-160,
//                     atomics on the run's accumulators), CSR of the kept matches, text lines (GC.cpp:1289-1322),
//                     active mask / classified level of the read for the next hierarchy level (GC.cpp:811-830).
// --fpr-query compares a libm double with the threshold.  The device evaluates the same expression with CUDA's
// lgamma/exp/pow (a few ulp from glibc's); a value within `fpr_band` of the threshold raises kFtAmbiguous and the host
// finishing stage redoes the level with libm -- decisions are therefore always the reference's.
// =====================================================================================================================
namespace
{
constexpr int      K4_THREADS = 256;
constexpr uint32_t kNoTuple = 0xFFFFFFFFu, kNoNode = 0xFFFFFFFFu, kFull = 0xFFFFFFFFu;

__global__ void k_tuple_starts(const uint64_t *__restrict__ tuples, uint64_t n, uint32_t *__restrict__ start)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const uint32_t r = (uint32_t)(tuples[i] >> kTupleReadShift);
    if (i == 0 || (uint32_t)(tuples[i - 1] >> kTupleReadShift) != r)
        start[r] = (uint32_t)i;
}

// LevelRt::lca2 (session.cpp): parent/depth walk; nodes outside the tree of the root meet at the root
__device__ __forceinline__ uint32_t dev_lca2(const int32_t *__restrict__ parent, const uint32_t *__restrict__ depth, int32_t root, uint32_t u, uint32_t v)
{
    int32_t a = (int32_t)u, b = (int32_t)v;
    while (a >= 0 && b >= 0 && a != b)
    {
        const uint32_t da = depth[a], db = depth[b];
        if (da > db)
            a = parent[a];
        else if (db > da)
            b = parent[b];
        else
        {
            a = parent[a];
            b = parent[b];
        }
    }
    return (a < 0 || b < 0) ? (uint32_t)root : (uint32_t)a;
}

// GC.cpp:498-501 + 588-601, same expression order as the host version (fpr_query_q, session.cpp)
__device__ double dev_fpr_query_q(uint32_t n_hashes, uint32_t count, double fpr)
{
    double       q  = 1;
    const double n  = (double)n_hashes;
    const double l0 = lgamma(n + 1);
    for (uint32_t i = 0; i <= count; i++)
    {
        const double k     = (double)i;
        const double binom = exp(l0 - lgamma(n - k + 1) - lgamma(k + 1));
        const double a     = __dmul_rn(binom, pow(fpr, k));
        q                  = __dsub_rn(q, __dmul_rn(a, pow(1 - fpr, (double)(n_hashes - i))));
    }
    return q;
}

// The value depends on (fpr of the target, n_hashes, count) only and a run sees few distinct triples: a direct-mapped,
// lossy cache in HBM shared by all batches of the level keeps the libm-style evaluation off the common path.  A slot is
// one aligned 16-byte word written and read with single 128-bit accesses.
__device__ __forceinline__ double dev_fpr_query_cached(const FinishParams &p, uint32_t node, uint32_t n_hashes, uint32_t count)
{
    // node: index into the level's [filter][node] tables (filter * n_nodes + node for levels with several filters)
    const unsigned long long key  = ((unsigned long long)p.node_class[node] << 32) | ((unsigned long long)(n_hashes & 0xFFFF) << 16) | (count & 0xFFFF);
    ulonglong2              *slot = reinterpret_cast<ulonglong2 *>(p.fpr_memo) + (splitmix64(key) & p.fpr_memo_mask);
    const ulonglong2         v    = __ldcg(slot);
    if (v.x == key)
        return __longlong_as_double((long long)v.y);
    const double q = dev_fpr_query_q(n_hashes, count, p.node_fpr[node]);
    __stcg(slot, make_ulonglong2(key, (unsigned long long)__double_as_longlong(q)));
    return q;
}

__device__ __forceinline__ uint32_t dec_digits(uint32_t v)
{
    return v < 10 ? 1 : v < 100 ? 2 : v < 1000 ? 3 : v < 10000 ? 4 : v < 100000 ? 5 : v < 1000000 ? 6 : v < 10000000 ? 7 : v < 100000000 ? 8 : v < 1000000000 ? 9 : 10;
}

// One thread per read (grid-stride): a read has few tuples, and what bounds the pass is the chain of dependent loads
// per read (record -> first tuple -> tuples -> fpr cache -> names), so the more reads in flight the better.
__global__ void __launch_bounds__(K4_THREADS) k_finish_select(const FinishParams p)
{
    const uint32_t lane   = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    unsigned long long acc = 0; // lane i of every warp accumulates totals[i]
    bool               ambiguous = false;
    // the loop bound is warp-uniform so that the warp reductions below see all lanes
    for (uint32_t r0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); r0 < p.n_reads; r0 += stride)
    {
        const uint32_t r     = r0 + lane;
        const bool     inb   = r < p.n_reads;
        const bool     act   = inb && (p.first ? true : p.active[r] != 0);
        FinishSizes    sz{0, 0, 0, 0};
        uint32_t       nacc = 0;
        uint2          one  = make_uint2(0, 0);
        uint32_t       c_small = 0, c_big = 0, c_proc = 0, c_len = 0, c_kmers = 0, c_class = 0, c_kmatch = 0, c_kclass = 0, c_match = 0, c_uniq = 0,
                 c_dfilter = 0, c_dfpr = 0, c_next = 0;
        if (act)
        {
            const uint32_t nh = p.n_hashes[r], l1 = p.len1[r], l2 = p.len2 ? p.len2[r] : 0;
            if (p.first)
            {
                if (l1 < p.w)
                    c_small = 1;
                else if (nh > 65535)
                    c_big = 1;
                else
                {
                    c_proc  = 1;
                    c_len   = l1 + l2;
                    c_kmers = nh;
                }
            }
            const uint32_t start = p.tuple_start[r];
            const uint32_t idl   = p.id_len[r];
            uint32_t       kept = 0, max_c = 0, min_c = nh, all_bytes = 0, one_node = kNoNode, one_cnt = 0;
            if (start != kNoTuple)
            {
                // ---- pass 1: sum runs of equal (read, node, filter), cap, cutoff; merge the filters of a node ----
                // Levels with several filters carry the filter index in the low `filter_bits` bits of the tuple's node field,
                // so the runs of one node come in --ibf order and the cross-filter merge of select_matches (GC.cpp:528-539)
                // is a walk over them: a filter's count replaces the node's only if it is strictly greater, and max / min of
                // the read are updated at EVERY such store -- a value overwritten later still lowers min (the reference's
                // behaviour, kept).
                const uint32_t fmask = (1u << p.filter_bits) - 1;
                uint64_t       run_key = ~0ull, sum = 0;
                bool           partial = false;
                uint32_t       cur_node = kNoNode, best = 0, best_f = 0;
                auto           emit    = [&]() {
                    if (cur_node != kNoNode && best)
                        p.entries[(uint64_t)start + nacc++] = ((uint64_t)cur_node << 32) | ((uint64_t)best_f << 24) | best;
                };
                auto flush = [&]() {
                    if (run_key == ~0ull)
                        return;
                    const uint32_t enc = (uint32_t)run_key & (kMaxNodes - 1), node = enc >> p.filter_bits, fi = enc & fmask;
                    if (node != cur_node)
                    {
                        emit();
                        cur_node = node;
                        best     = 0;
                    }
                    const uint32_t cutoff = threshold_cutoff(nh, p.rel_cutoffs[fi]);
                    bool           ok     = true;
                    if (p.is_hibf)
                    {
                        if (partial)
                            sum &= 0xFFFF;
                        if (sum < cutoff || sum == 0)
                            ok = false;
                        if (sum > nh)
                            sum = nh;
                    }
                    else if (partial)
                    {
                        if (sum > nh)
                            sum = nh;
                        if (sum < cutoff)
                            ok = false;
                    }
                    if (ok && (uint32_t)sum > best)
                    {
                        best   = (uint32_t)sum;
                        best_f = fi;
                        max_c  = max(max_c, best);
                        min_c  = min(min_c, best);
                    }
                };
                for (uint64_t i = start; i < p.n_tuples; ++i)
                {
                    const uint64_t t = p.tuples[i];
                    if ((uint32_t)(t >> kTupleReadShift) != r)
                        break;
                    const uint64_t key = t >> kTupleNodeShift; // (read, node)
                    if (key != run_key)
                    {
                        flush();
                        run_key = key;
                        sum     = 0;
                        partial = false;
                    }
                    sum += t & 0xFFFF;
                    partial |= ((t >> 16) & 1) != 0;
                }
                flush();
                emit();
                // ---- pass 2: rel-filter, fpr-query, LCA, sizes ----
                if (nacc)
                {
                    const uint64_t thr_ceil         = (uint64_t)ceil(__dmul_rn((double)(max_c - min_c), p.rel_filter));
                    const double   threshold_filter = (double)((uint64_t)max_c - thr_ceil);
                    for (uint32_t j = 0; j < nacc; ++j)
                    {
                        const uint64_t e    = p.entries[(uint64_t)start + j];
                        const uint32_t node = (uint32_t)(e >> 32), cnt = (uint32_t)(e & 0xFFFF), fi = (uint32_t)(e >> 24) & 15;
                        uint32_t       status = 0;
                        if ((double)cnt >= threshold_filter)
                        {
                            if (p.fpr_query < 1.0)
                            {
                                const double q = dev_fpr_query_cached(p, fi * p.n_nodes + node, nh, cnt);
                                if (fabs(q - p.fpr_query) <= p.fpr_band)
                                    ambiguous = true;
                                if (q > p.fpr_query)
                                    status = 2;
                            }
                        }
                        else
                            status = 1;
                        if (status)
                            p.entries[(uint64_t)start + j] = e | ((uint64_t)status << 16);
                        if (status == 1)
                            ++c_dfilter;
                        else if (status == 2)
                            ++c_dfpr;
                        else
                        {
                            if (p.output_all)
                                all_bytes += idl + 1 + (p.name_off[node + 1] - p.name_off[node]) + 1 + dec_digits(cnt) + 1;
                            if (kept == 0)
                            {
                                one_node = node;
                                one_cnt  = cnt;
                            }
                            else if (!p.skip_lca)
                                one_node = dev_lca2(p.parent, p.depth, p.root, one_node, node);
                            ++kept;
                        }
                    }
                    c_match = kept;
                }
            }
            if (kept > 0)
            {
                c_class  = 1;
                c_kclass = nh;
                c_kmatch = max_c;
                if (kept == 1)
                    c_uniq = 1;
                else
                {
                    if (p.skip_lca)
                        one_node = (uint32_t)p.root;
                    one_cnt = max_c;
                }
                one          = make_uint2(one_node, one_cnt);
                sz.kept      = kept;
                sz.all_bytes = all_bytes;
                if (!p.skip_lca && p.output_lca)
                    sz.one_bytes = idl + 1 + (p.name_off[one_node + 1] - p.name_off[one_node]) + 1 + dec_digits(one_cnt) + 1;
            }
            else
            {
                if (p.last && p.output_unc)
                    sz.unc_bytes = idl + 1;
                c_next = nh <= 65535 ? nh : 0;
            }
        }
        if (inb)
        {
            p.sizes[r] = sz;
            p.n_acc[r] = nacc;
            p.one[r]   = one;
        }
#define K4_ADD(idx, v)                                         \
    do                                                         \
    {                                                          \
        const uint32_t s_ = __reduce_add_sync(kFull, (v));     \
        if (lane == (idx))                                     \
            acc += s_;                                         \
    } while (0)
        if (p.first)
        {
            K4_ADD(kFtSkippedSmall, c_small);
            K4_ADD(kFtSkippedBig, c_big);
            K4_ADD(kFtProcessed, c_proc);
            K4_ADD(kFtLength, c_len);
            K4_ADD(kFtKmers, c_kmers);
        }
        K4_ADD(kFtClassified, c_class);
        K4_ADD(kFtKmersMatches, c_kmatch);
        K4_ADD(kFtKmersClassified, c_kclass);
        K4_ADD(kFtMatches, c_match);
        K4_ADD(kFtUnique, c_uniq);
        K4_ADD(kFtDiscFilter, c_dfilter);
        K4_ADD(kFtDiscFpr, c_dfpr);
        K4_ADD(kFtActiveHashesNext, c_next);
#undef K4_ADD
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        p.sizes[p.n_reads] = FinishSizes{0, 0, 0, 0};
    if (__any_sync(kFull, ambiguous) && lane == kFtAmbiguous)
        acc += 1;
    __shared__ unsigned long long s_acc[kFinishTotals];
    if (threadIdx.x < kFinishTotals)
        s_acc[threadIdx.x] = 0;
    __syncthreads();
    if (lane < kFinishTotals && acc)
        atomicAdd(&s_acc[lane], acc);
    __syncthreads();
    if (threadIdx.x < kFinishTotals && s_acc[threadIdx.x])
        atomicAdd(&p.totals[threadIdx.x], s_acc[threadIdx.x]);
}

struct FinishSizesAdd
{
    __host__ __device__ __forceinline__ FinishSizes operator()(const FinishSizes &a, const FinishSizes &b) const
    {
        return FinishSizes{a.kept + b.kept, a.all_bytes + b.all_bytes, a.one_bytes + b.one_bytes, a.unc_bytes + b.unc_bytes};
    }
};

__device__ __forceinline__ char *put_line(char *o, const uint8_t *id, uint32_t idl, const char *name, uint32_t nl, uint32_t v)
{
    for (uint32_t i = 0; i < idl; ++i)
        *o++ = (char)id[i];
    *o++ = '\t';
    for (uint32_t i = 0; i < nl; ++i)
        *o++ = name[i];
    *o++ = '\t';
    const uint32_t d = dec_digits(v);
    for (uint32_t i = d; i-- > 0;)
    {
        o[i] = (char)('0' + v % 10);
        v /= 10;
    }
    o += d;
    *o++ = '\n';
    return o;
}

__global__ void __launch_bounds__(K4_THREADS) k_finish_write(const FinishParams p)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    if (blockIdx.x == 0 && threadIdx.x == 0)
        p.match_off[p.n_reads] = p.offs[p.n_reads].kept;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < p.n_reads; r += stride)
    {
        const FinishSizes o = p.offs[r];
        p.match_off[r]      = o.kept;
        const bool act      = p.first ? true : p.active[r] != 0;
        if (!act)
            continue;
        const uint32_t nacc  = p.n_acc[r];
        const uint32_t start = nacc ? p.tuple_start[r] : 0;
        const uint8_t *id    = p.blk1 + p.id_off[r];
        const uint32_t idl   = p.id_len[r];
        uint32_t       kept  = 0;
        char          *out   = p.all_text + o.all_bytes;
        for (uint32_t j = 0; j < nacc; ++j)
        {
            const uint64_t e    = p.entries[(uint64_t)start + j];
            const uint32_t node = (uint32_t)(e >> 32), cnt = (uint32_t)(e & 0xFFFF), status = (uint32_t)(e >> 16) & 3;
            atomicAdd(&p.rep[(uint64_t)node * 5 + (status == 0 ? 0 : status == 1 ? 3 : 4)], 1ull);
            if (status)
                continue;
            p.match_target[o.kept + kept] = node;
            p.match_count[o.kept + kept]  = cnt;
            ++kept;
            if (p.output_all)
                out = put_line(out, id, idl, p.names + p.name_off[node], p.name_off[node + 1] - p.name_off[node], cnt);
        }
        if (kept > 0)
        {
            const uint2 one = p.one[r];
            atomicAdd(&p.rep[(uint64_t)one.x * 5 + (kept == 1 ? 2 : 1)], 1ull);
            if (!p.skip_lca && p.output_lca)
                put_line(p.one_text + o.one_bytes, id, idl, p.names + p.name_off[one.x], p.name_off[one.x + 1] - p.name_off[one.x], one.y);
            p.active[r]     = 0;
            p.read_level[r] = (uint8_t)p.level;
        }
        else
        {
            p.active[r] = 1;
            if (p.last && p.output_unc)
            {
                char *u = p.unc_text + o.unc_bytes;
                for (uint32_t i = 0; i < idl; ++i)
                    u[i] = (char)id[i];
                u[idl] = '\n';
            }
        }
    }
}
} // namespace

void launch_finish_select(const FinishParams &p, cudaStream_t st)
{
    cudaMemsetAsync(p.tuple_start, 0xFF, (size_t)p.n_reads * 4, st);
    cudaMemsetAsync(p.totals, 0, kFinishTotals * 8, st);
    if (p.n_tuples)
        k_tuple_starts<<<(unsigned)((p.n_tuples + 255) / 256), 256, 0, st>>>(p.tuples, p.n_tuples, p.tuple_start);
    const unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)p.n_reads + K4_THREADS - 1) / K4_THREADS, 148u * 8);
    k_finish_select<<<std::max(1u, blocks), K4_THREADS, 0, st>>>(p);
}

size_t finish_scan_tmp_bytes(uint32_t n_reads)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveScan(nullptr, bytes, (const FinishSizes *)nullptr, (FinishSizes *)nullptr, FinishSizesAdd{}, FinishSizes{0, 0, 0, 0}, (int)(n_reads + 1));
    return bytes + 256;
}

void launch_finish_scan(const FinishParams &p, void *tmp, size_t tmp_bytes, cudaStream_t st)
{
    cub::DeviceScan::ExclusiveScan(tmp, tmp_bytes, (const FinishSizes *)p.sizes, p.offs, FinishSizesAdd{}, FinishSizes{0, 0, 0, 0}, (int)(p.n_reads + 1), st);
}

void launch_finish_write(const FinishParams &p, cudaStream_t st)
{
    const unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)p.n_reads + K4_THREADS - 1) / K4_THREADS, 148u * 8);
    k_finish_write<<<std::max(1u, blocks), K4_THREADS, 0, st>>>(p);
}

// =====================================================================================================================
// EM reassignment (src/ganon/reassign.py) on the matches of the whole run kept in HBM.
// Probabilities of one iteration share a denominator (reassign.py:106-107, 125-128), so "highest probability" is
// "highest integer weight": the device compares the counts themselves; the host keeps the double-precision
// probabilities only for the convergence test (|old - new| summed in the reference's target order).
// =====================================================================================================================
namespace
{
__global__ void k_em_sizes(const FinishSizes *__restrict__ sizes, const uint32_t *__restrict__ id_len, uint32_t n, EmSizes *__restrict__ es)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n)
        return;
    const bool cls = r < n && sizes[r].kept != 0;
    es[r]          = EmSizes{cls ? 1ull : 0ull, cls ? (unsigned long long)id_len[r] : 0ull};
}
struct EmSizesAdd
{
    __host__ __device__ __forceinline__ EmSizes operator()(const EmSizes &a, const EmSizes &b) const { return EmSizes{a.reads + b.reads, a.id_bytes + b.id_bytes}; }
};
__global__ void k_em_append_reads(const FinishSizes *__restrict__ sizes, const EmSizes *__restrict__ offs, const uint64_t *__restrict__ match_off,
                                  const uint32_t *__restrict__ id_off, const uint32_t *__restrict__ id_len, const uint8_t *__restrict__ blk, uint32_t n, EmStoreDev st,
                                  uint64_t base_reads, uint64_t base_matches, uint64_t base_ids)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0)
    { // sentinels of the grown store
        st.off[base_reads + offs[n].reads]    = base_matches + match_off[n];
        st.id_off[base_reads + offs[n].reads] = base_ids + offs[n].id_bytes;
    }
    if (r >= n || sizes[r].kept == 0)
        return;
    const uint64_t j = base_reads + offs[r].reads;
    st.off[j]        = base_matches + match_off[r];
    const uint64_t o = base_ids + offs[r].id_bytes;
    st.id_off[j]     = o;
    const uint8_t *id = blk + id_off[r];
    for (uint32_t i = 0; i < id_len[r]; ++i)
        st.ids[o + i] = (char)id[i];
}
__global__ void k_em_append_matches(const uint32_t *__restrict__ mt, const uint32_t *__restrict__ mc, uint64_t n, const uint32_t *__restrict__ map, uint32_t *__restrict__ dt,
                                    uint32_t *__restrict__ dc)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    dt[i] = map[mt[i]];
    dc[i] = mc[i];
}
__global__ void k_em_first_pos(const uint32_t *__restrict__ tgt, uint64_t n, unsigned long long *__restrict__ first_pos)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n)
        atomicMin(&first_pos[tgt[i]], (unsigned long long)i);
}
__global__ void k_em_initial(EmStoreDev st, uint64_t n_reads, unsigned long long *__restrict__ initial)
{
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_reads)
        return;
    const uint64_t a = st.off[r];
    if (st.off[r + 1] - a == 1)
        atomicAdd(&initial[st.tgt[a]], 1ull);
}
// get_top_match (reassign.py:229-241): first match unless a later one has a strictly higher, non-zero weight
__device__ __forceinline__ uint64_t em_top(const EmStoreDev &st, uint64_t a, uint64_t b, const unsigned long long *__restrict__ weight)
{
    uint64_t           best = a;
    unsigned long long bw   = 0;
    for (uint64_t i = a; i < b; ++i)
    {
        const unsigned long long w = weight[st.tgt[i]];
        if (w > bw)
        {
            bw   = w;
            best = i;
        }
    }
    return best;
}
__global__ void k_em_assign(EmStoreDev st, uint64_t n_reads, const unsigned long long *__restrict__ weight, unsigned long long *__restrict__ counts)
{
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_reads)
        return;
    const uint64_t a = st.off[r], b = st.off[r + 1];
    if (b - a > 1)
        atomicAdd(&counts[st.tgt[em_top(st, a, b, weight)]], 1ull);
}
__global__ void k_em_one(EmStoreDev st, uint64_t n_reads, const unsigned long long *__restrict__ weight, const uint32_t *__restrict__ name_off, const char *__restrict__ names,
                         uint64_t *__restrict__ line_len, const uint64_t *__restrict__ line_off, char *__restrict__ out, unsigned long long *__restrict__ n_multi)
{
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_reads)
    {
        if (r == n_reads && out == nullptr)
            line_len[r] = 0;
        return;
    }
    const uint64_t a = st.off[r], b = st.off[r + 1];
    const uint64_t m = b - a == 1 ? a : em_top(st, a, b, weight);
    const uint32_t t = st.tgt[m], k = st.cnt[m];
    const uint32_t idl = (uint32_t)(st.id_off[r + 1] - st.id_off[r]), nl = name_off[t + 1] - name_off[t];
    if (out == nullptr)
    {
        line_len[r] = idl + 1 + nl + 1 + dec_digits(k) + 1;
        if (b - a > 1)
            atomicAdd(n_multi, 1ull);
        return;
    }
    put_line(out + line_off[r], reinterpret_cast<const uint8_t *>(st.ids + st.id_off[r]), idl, names + name_off[t], nl, k);
}
} // namespace

size_t em_scan_tmp_bytes(uint32_t n_reads)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveScan(nullptr, bytes, (const EmSizes *)nullptr, (EmSizes *)nullptr, EmSizesAdd{}, EmSizes{0, 0}, (int)(n_reads + 1));
    return bytes + 256;
}
void launch_em_sizes(const FinishSizes *sizes, const uint32_t *id_len, uint32_t n_reads, EmSizes *es, EmSizes *offs, void *tmp, size_t tmp_bytes, cudaStream_t st)
{
    k_em_sizes<<<(n_reads + 1 + 255) / 256, 256, 0, st>>>(sizes, id_len, n_reads, es);
    cub::DeviceScan::ExclusiveScan(tmp, tmp_bytes, (const EmSizes *)es, offs, EmSizesAdd{}, EmSizes{0, 0}, (int)(n_reads + 1), st);
}
void launch_em_append(const FinishSizes *sizes, const EmSizes *offs, const uint64_t *match_off, const uint32_t *match_target, const uint32_t *match_count,
                      uint64_t n_matches, const uint32_t *id_off, const uint32_t *id_len, const uint8_t *blk, uint32_t n_reads, const uint32_t *node_to_target,
                      EmStoreDev store, uint64_t base_reads, uint64_t base_matches, uint64_t base_ids, cudaStream_t st)
{
    k_em_append_reads<<<(n_reads + 255) / 256, 256, 0, st>>>(sizes, offs, match_off, id_off, id_len, blk, n_reads, store, base_reads, base_matches, base_ids);
    if (n_matches)
        k_em_append_matches<<<(unsigned)((n_matches + 255) / 256), 256, 0, st>>>(match_target, match_count, n_matches, node_to_target, store.tgt + base_matches,
                                                                                    store.cnt + base_matches);
}
void launch_em_first_pos(EmStoreDev store, uint64_t n_matches, unsigned long long *first_pos, cudaStream_t st)
{
    if (n_matches)
        k_em_first_pos<<<(unsigned)((n_matches + 255) / 256), 256, 0, st>>>(store.tgt, n_matches, first_pos);
}
void launch_em_initial(EmStoreDev store, uint64_t n_reads, unsigned long long *initial, cudaStream_t st)
{
    if (n_reads)
        k_em_initial<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(store, n_reads, initial);
}
void launch_em_assign(EmStoreDev store, uint64_t n_reads, const unsigned long long *weight, unsigned long long *counts, cudaStream_t st)
{
    if (n_reads)
        k_em_assign<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(store, n_reads, weight, counts);
}
void launch_em_one(EmStoreDev store, uint64_t n_reads, const unsigned long long *weight, const uint32_t *name_off, const char *names, uint64_t *line_len,
                   const uint64_t *line_off, char *out, unsigned long long *n_multi, cudaStream_t st)
{
    k_em_one<<<(unsigned)((n_reads + 1 + 255) / 256), 256, 0, st>>>(store, n_reads, weight, name_off, names, line_len, line_off, out, n_multi);
}
size_t em_scan64_tmp_bytes(uint64_t n)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int64_t)n);
    return bytes + 256;
}
void launch_scan64(const uint64_t *in, uint64_t *out, uint64_t n, void *tmp, size_t tmp_bytes, cudaStream_t st)
{
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int64_t)n, st);
}

// Equal read ids in the store (reassign.py keys its matches by read id: reads that share one are a single read to it).
// Every id is hashed (FNV-1a, 64 bits, finished with a multiply-shift mix), the hashes are sorted, equal neighbours are
// counted: zero means that no two ids are equal; anything else sends the store through the exact regrouping on the host
// (em_merge.cpp), which compares the bytes.
namespace
{
__global__ void k_em_hash_ids(EmStoreDev st, uint64_t n_reads, uint64_t *__restrict__ hash)
{
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_reads)
        return;
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t i = st.id_off[r], e = st.id_off[r + 1]; i < e; ++i)
        h = (h ^ (uint8_t)st.ids[i]) * 0x100000001b3ull;
    h ^= h >> 32;
    h *= 0x9e3779b97f4a7c15ull;
    hash[r] = h ^ (h >> 29);
}
__global__ void k_em_equal_neighbours(const uint64_t *__restrict__ sorted, uint64_t n, unsigned long long *__restrict__ n_equal)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const bool     eq = i + 1 < n && sorted[i] == sorted[i + 1];
    const unsigned m  = __ballot_sync(0xffffffffu, eq);
    if ((threadIdx.x & 31) == 0 && m)
        atomicAdd(n_equal, (unsigned long long)__popc(m));
}
} // namespace
size_t em_equal_ids_tmp_bytes(uint64_t n_reads)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int64_t)n_reads);
    return bytes + 256;
}
void launch_em_equal_ids(EmStoreDev store, uint64_t n_reads, uint64_t *keys_a, uint64_t *keys_b, void *tmp, size_t tmp_bytes, unsigned long long *n_equal,
                         cudaStream_t st)
{
    cudaMemsetAsync(n_equal, 0, sizeof(unsigned long long), st);
    if (n_reads < 2)
        return;
    const unsigned blocks = (unsigned)((n_reads + 255) / 256);
    k_em_hash_ids<<<blocks, 256, 0, st>>>(store, n_reads, keys_a);
    cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys_a, keys_b, (int64_t)n_reads, 0, 64, st);
    k_em_equal_neighbours<<<blocks, 256, 0, st>>>(keys_b, n_reads, n_equal);
}

// =====================================================================================================================
// build-side helpers
// =====================================================================================================================
namespace
{
__global__ void k_fill_random(uint64_t *__restrict__ data, uint64_t rows, uint32_t row_words, uint32_t w0, uint32_t total_words, uint64_t bins,
                              uint64_t seed, int and_terms)
{
    // word (row, w) of the whole filter has index row*total_words + w; a shard holds columns [w0, w0+row_words)
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n      = rows * row_words;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
    {
        const uint64_t row = j / row_words, w = w0 + (j - row * row_words);
        const uint64_t i   = row * total_words + w;
        uint64_t       v   = ~0ULL;
        for (int t = 0; t < and_terms; ++t)
            v &= splitmix64(seed + i * 8 + (uint64_t)t);
        // padding bins [bins, 64*total_words) stay zero (IBF.hpp:238-240)
        const uint64_t lo = w * 64;
        if (lo + 64 > bins)
            v &= (lo >= bins) ? 0ULL : ((1ULL << (bins - lo)) - 1);
        data[j] = v;
    }
}

__global__ void k_emplace(uint64_t *__restrict__ data, uint64_t bin_size, uint32_t hash_shift, uint32_t hash_funs, uint32_t row_words, uint32_t w0,
                          const uint64_t *__restrict__ hashes, const uint32_t *__restrict__ bins, uint64_t n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * hash_funs)
        return;
    const uint64_t m  = i / hash_funs;
    const uint32_t fn = (uint32_t)(i - m * hash_funs);
    const uint32_t b  = bins[m];
    const uint32_t w  = b >> 6;
    if (w < w0 || w >= w0 + row_words) // bin lives in another shard
        return;
    const uint64_t row = ibf_row(hashes[m], ibf_seed(fn), hash_shift, bin_size);
    atomicOr((unsigned long long *)(data + row * row_words + (w - w0)), 1ULL << (b & 63));
}
} // namespace

void launch_fill_random(uint64_t *data, uint64_t rows, uint32_t row_words, uint32_t w0, uint32_t total_words, uint64_t bins, uint64_t seed,
                        int and_terms, cudaStream_t st)
{
    k_fill_random<<<148 * 16, 256, 0, st>>>(data, rows, row_words, w0, total_words, bins, seed, and_terms);
}

void launch_emplace(uint64_t *data, uint64_t bin_size, uint32_t hash_shift, uint32_t hash_funs, uint32_t row_words, uint32_t w0,
                    const uint64_t *hashes, const uint32_t *bins, uint64_t n, cudaStream_t st)
{
    if (n == 0)
        return;
    const uint64_t threads = n * hash_funs;
    k_emplace<<<(uint32_t)((threads + 255) / 256), 256, 0, st>>>(data, bin_size, hash_shift, hash_funs, row_words, w0, hashes, bins, n);
}

// =====================================================================================================================
// K1 -- FASTQ record index (strict 4-line records; anything else is indexed by the host reader in reads.cpp)
// =====================================================================================================================
namespace
{
constexpr int K1_THREADS = 256;
constexpr int K1_BYTES_PER_THREAD = 16;
constexpr int K1_TILE = K1_THREADS * K1_BYTES_PER_THREAD; // 4 KiB per CTA

__device__ __forceinline__ uint32_t newline_mask16(const uint8_t *blk, uint64_t pos, uint64_t n_bytes)
{
    // bit i set: blk[pos+i] == '\n'
    uint32_t m = 0;
    if (pos + 16 <= n_bytes && ((uintptr_t)(blk + pos) & 15) == 0)
    {
        const uint4 v = *reinterpret_cast<const uint4 *>(blk + pos);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                m |= (((w[q] >> (8 * b)) & 0xffu) == '\n') ? (1u << (q * 4 + b)) : 0u;
    }
    else
    {
        for (int i = 0; i < 16; ++i)
            if (pos + i < n_bytes && blk[pos + i] == '\n')
                m |= 1u << i;
    }
    return m;
}

__global__ void __launch_bounds__(K1_THREADS) k_count_newlines(const uint8_t *__restrict__ blk, uint64_t n_bytes, uint32_t *__restrict__ tile_counts)
{
    const uint64_t pos = (uint64_t)blockIdx.x * K1_TILE + (uint64_t)threadIdx.x * K1_BYTES_PER_THREAD;
    uint32_t       c   = pos < n_bytes ? __popc(newline_mask16(blk, pos, n_bytes)) : 0;
    __shared__ uint32_t s[K1_THREADS / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        c += __shfl_down_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0)
        s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t t = 0;
        for (int i = 0; i < K1_THREADS / 32; ++i)
            t += s[i];
        tile_counts[blockIdx.x] = t;
    }
}

// line_start[j] = byte offset of line j (line 0 starts at 0); line_start[n_lines] = one past the last newline
__global__ void __launch_bounds__(K1_THREADS)
    k_line_starts(const uint8_t *__restrict__ blk, uint64_t n_bytes, const uint32_t *__restrict__ tile_base, uint32_t *__restrict__ line_start, uint32_t cap_lines)
{
    const uint64_t pos = (uint64_t)blockIdx.x * K1_TILE + (uint64_t)threadIdx.x * K1_BYTES_PER_THREAD;
    const uint32_t m   = pos < n_bytes ? newline_mask16(blk, pos, n_bytes) : 0;
    const uint32_t c   = __popc(m);
    // block exclusive scan of c
    __shared__ uint32_t s[K1_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t       incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += y;
    }
    if (lane == 31)
        s[wid] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t i = 0; i < wid; ++i)
        wbase += s[i];
    uint32_t j = tile_base[blockIdx.x] + wbase + incl - c; // newlines before this thread's bytes
    if (blockIdx.x == 0 && threadIdx.x == 0)
        line_start[0] = 0;
    uint32_t mm = m;
    while (mm)
    {
        const uint32_t b = __ffs(mm) - 1;
        mm &= mm - 1;
        ++j; // newline number j-1 ends line j-1; line j starts right after it
        if (j < cap_lines)
            line_start[j] = (uint32_t)(pos + b + 1);
    }
}

// one thread per record: derive spans, check the '@' / '+' markers and |seq| == |qual| (format_fastq.hpp:124-262)
__global__ void k_fastq_records(const uint8_t *__restrict__ blk, const uint32_t *__restrict__ line_start, uint32_t n_records, FastqIndexOut out)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_records)
        return;
    const uint32_t l0 = line_start[4 * r], l1 = line_start[4 * r + 1], l2 = line_start[4 * r + 2], l3 = line_start[4 * r + 3],
                   l4 = line_start[4 * r + 4];
    const uint32_t seq_len = l2 - 1 - l1, qual_len = l4 - 1 - l3;
    out.id_off[r]  = l0 + 1;
    out.id_len[r]  = l1 - 1 - (l0 + 1);
    out.seq_off[r] = l1;
    out.seq_len[r] = seq_len;
    bool ok = blk[l0] == '@' && blk[l2] == '+' && seq_len == qual_len;
    // a '+' inside the sequence line, blanks, or illegal letters are caught by the character check below
    if (!ok)
    {
        atomicAdd(&out.status[1], 1u);
        atomicMin(&out.status[2], r);
    }
}

// Unwrapped FASTA (">id\nSEQUENCE\n" per record, what read files in FASTA format look like): spans from line_start[2r..2r+2].
// A sequence line that starts like a header, or a header line that does not, makes the block irregular (wrapped sequences,
// ';' comment lines, blanks: the host reader reproduces seqan3's format_fasta there).
__global__ void k_fasta_records(const uint8_t *__restrict__ blk, uint64_t n_bytes, const uint32_t *__restrict__ line_start, uint32_t n_records, FastqIndexOut out)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_records)
        return;
    const uint32_t l0 = line_start[2 * r], l1 = line_start[2 * r + 1], l2 = line_start[2 * r + 2];
    out.id_off[r]  = l0 + 1;
    out.id_len[r]  = l1 - 1 - (l0 + 1);
    out.seq_off[r] = l1;
    out.seq_len[r] = l2 - 1 - l1;
    // the line after the sequence line must open the next record (or be the end of the block): otherwise the sequence is
    // wrapped over several lines and this is not the 2-line form
    // (seqan3 skips blanks between '>' and the id: such headers go to the host reader as well)
    const bool ok = blk[l0] == '>' && blk[l0 + 1] != ' ' && blk[l0 + 1] != '\t' && l2 - 1 > l1 && blk[l1] != '>' && blk[l1] != ';' && (l2 >= n_bytes || blk[l2] == '>');
    if (!ok)
    {
        atomicAdd(&out.status[1], 1u);
        atomicMin(&out.status[2], r);
    }
}

// warp per record: every sequence character must be legal for dna15 (parse_error otherwise)
__global__ void k_validate_seq(const uint8_t *__restrict__ blk, const uint32_t *__restrict__ seq_off, const uint32_t *__restrict__ seq_len,
                               uint32_t n_records, uint32_t *__restrict__ status)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_records)
        return;
    const uint8_t *s = blk + seq_off[r];
    const uint32_t L = seq_len[r];
    bool bad = false;
    for (uint32_t i = lane; i < L; i += 32)
        bad |= !dna15_valid(s[i]);
    if (__any_sync(0xffffffffu, bad) && lane == 0)
    {
        atomicAdd(&status[3], 1u);
        atomicMin(&status[2], r);
    }
}
} // namespace

size_t fastq_index_tmp_bytes(uint64_t n_bytes)
{
    const uint64_t tiles = (n_bytes + K1_TILE - 1) / K1_TILE + 1;
    size_t         scan  = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)tiles);
    return 2 * (tiles + 64) * sizeof(uint32_t) + scan + 512;
}

// Phase 1a: count the newlines of the block; *n_lines_dev = total.  tmp keeps the per-tile bases for phase 1b.
void launch_fastq_count(const uint8_t *blk, uint64_t n_bytes, uint32_t *n_lines_dev, void *tmp, size_t tmp_bytes, cudaStream_t st)
{
    const uint32_t tiles = (uint32_t)((n_bytes + K1_TILE - 1) / K1_TILE);
    uint32_t *tile_counts = reinterpret_cast<uint32_t *>(tmp);
    uint32_t *tile_base   = tile_counts + (tiles + 64);
    char     *scan_tmp    = reinterpret_cast<char *>(tile_base + (tiles + 64));
    uintptr_t a           = ((uintptr_t)scan_tmp + 255) & ~(uintptr_t)255;
    size_t    scan_bytes  = tmp_bytes - (size_t)(a - (uintptr_t)tmp);
    cudaMemsetAsync(tile_counts + tiles, 0, sizeof(uint32_t), st);
    if (tiles)
        k_count_newlines<<<tiles, K1_THREADS, 0, st>>>(blk, n_bytes, tile_counts);
    cub::DeviceScan::ExclusiveSum((void *)a, scan_bytes, tile_counts, tile_base, (int)(tiles + 1), st);
    cudaMemcpyAsync(n_lines_dev, tile_base + tiles, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st);
}

// Phase 1b: line_start[0..min(n_lines, cap_lines-1)] from the tile bases left in tmp by phase 1a.
void launch_fastq_line_starts(const uint8_t *blk, uint64_t n_bytes, uint32_t *line_start, uint32_t cap_lines, void *tmp, cudaStream_t st)
{
    const uint32_t tiles = (uint32_t)((n_bytes + K1_TILE - 1) / K1_TILE);
    uint32_t *tile_counts = reinterpret_cast<uint32_t *>(tmp);
    uint32_t *tile_base   = tile_counts + (tiles + 64);
    if (tiles)
        k_line_starts<<<tiles, K1_THREADS, 0, st>>>(blk, n_bytes, tile_base, line_start, cap_lines);
    else
        cudaMemsetAsync(line_start, 0, sizeof(uint32_t), st);
}

// Phase 2: record spans + marker / length / alphabet validation.  out.status must be {0, 0, 0xffffffff, 0} on entry.
void launch_fasta_records(const uint8_t *blk, uint64_t n_bytes, const uint32_t *line_start, uint32_t n_records, FastqIndexOut out, cudaStream_t st)
{
    if (n_records == 0)
        return;
    k_fasta_records<<<(n_records + 255) / 256, 256, 0, st>>>(blk, n_bytes, line_start, n_records, out);
    const uint64_t threads = (uint64_t)n_records * 32;
    k_validate_seq<<<(uint32_t)((threads + 255) / 256), 256, 0, st>>>(blk, out.seq_off, out.seq_len, n_records, out.status);
}

void launch_fastq_records(const uint8_t *blk, const uint32_t *line_start, uint32_t n_records, FastqIndexOut out, cudaStream_t st)
{
    if (n_records == 0)
        return;
    k_fastq_records<<<(n_records + 255) / 256, 256, 0, st>>>(blk, line_start, n_records, out);
    const uint64_t threads = (uint64_t)n_records * 32;
    k_validate_seq<<<(uint32_t)((threads + 255) / 256), 256, 0, st>>>(blk, out.seq_off, out.seq_len, n_records, out.status);
}

} // namespace gnb
