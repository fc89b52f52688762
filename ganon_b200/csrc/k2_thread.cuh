// K2t -- minimisers, one THREAD per read, or per segment of a long read (k <= 29, w - k + 1 <= 32).
//
// The warp-per-read kernel (k_minimisers) spends ~1350 warp-instructions per 150 bp read, half of them on the doubled
// position-tagged minima that emulate the reference's sliding window with parallel scans (profiles/r01_ncu_summary_k2_k4_k3h.md).
// A thread that walks its own read can run the reference's state machine itself (minimiser.hpp:421-472):
//     first window, or the tracked minimiser left the window -> rightmost minimum of the window, emitted unconditionally;
//     else a strictly smaller value entered                   -> it becomes the minimiser, emitted;
//     else                                                    -> nothing.
// A literal rescan of the window would make every lane of a warp wait for the lane that rescans (with 32 lanes some lane
// does at nearly every position), so the window minimum is kept in a form with uniform control flow instead: the values are
// cut into blocks of W = w - k + 1; with suffix minima of the previous block and the running prefix minimum of the current
// one, the minimum of any window is one comparison (van Herk / Gil-Werman).  All lanes are at the same position of their
// reads, so block ends (where the ring of W values in shared memory is turned into suffix minima) coincide: no divergence
// apart from read lengths.  "Rightmost on ties" is carried by a 5-bit tag below the value: key = value << 5 | (W-1-offset in
// block), so that a plain unsigned minimum prefers the later of two equal values; across the two blocks the current
// (right) block wins ties; 2k + 5 <= 63 bounds k at 29, larger k stay on the warp kernel.  (sm_100a has no 64-bit integer
// or FP64 min instruction -- fmin() on the same bits compiles to DSETP + selects + NaN fix-up, 5 instructions -- so a
// minimum is two compares and two selects.)
// Bases come in through aligned 8-byte loads, fetched one word ahead and re-aligned with PRMT, and are decoded by a
// 256-entry table in shared memory that also carries the complement already shifted to its place in the
// reverse-complement k-mer.  Shared memory is addressed through 32-bit shared-space addresses (ld/st.shared).
//
// The per-thread core below is plain C++ (no warp collectives): tests/native/k2t_host.cpp compiles it with g++ and the CPU
// suite checks it against the oracle (tests/test_k2t_cpu.py); the kernel wrapper is in kernels.cu.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define K2T_D __device__ __forceinline__
#define K2T_M __device__ __forceinline__
#else
#define K2T_D static inline
#define K2T_M inline
#endif

namespace k2t
{

constexpr int      kThreads  = 128; // threads (= reads in flight) per CTA
constexpr uint32_t kTagBits  = 5;
constexpr uint32_t kTagMask  = (1u << kTagBits) - 1;
constexpr uint32_t kMaxW     = 1u << kTagBits;      // window of at most 32 values
constexpr uint32_t kMaxK     = (63 - kTagBits) / 2; // 29
constexpr uint64_t kBigKey   = ~0ull;               // above every key
constexpr uint32_t kLutBytes = 256 * 8;

// Table entry of a sequence character: rank = seqan3 dna4 char_to_rank incl. the IUPAC conversion (dna4.hpp:166-205:
// c y s b -> 1, g k -> 2, t u -> 3, anything else 0), comp = the complement (3 - rank, dna4.hpp:95-98) shifted to bit
// (2(k-1)) mod 32 -- its place in the low or the high word of the reverse-complement k-mer.
struct LutEntry
{
    uint32_t rank, comp;
};

K2T_D uint32_t rank_of(uint32_t c)
{
    constexpr uint64_t T = (1ull << (2 * ('c' - 'a'))) | (1ull << (2 * ('y' - 'a'))) | (1ull << (2 * ('s' - 'a'))) | (1ull << (2 * ('b' - 'a'))) |
                           (2ull << (2 * ('g' - 'a'))) | (2ull << (2 * ('k' - 'a'))) | (3ull << (2 * ('t' - 'a'))) | (3ull << (2 * ('u' - 'a')));
    const uint32_t i  = (c | 0x20u) - 'a';
    const uint32_t lo = (uint32_t)T, hi = (uint32_t)(T >> 32);
    const uint32_t w  = i < 16 ? lo : hi;
    return i < 26 ? (w >> ((i & 15) * 2)) & 3u : 0u;
}

K2T_D LutEntry lut_entry(uint32_t c, uint32_t k)
{
    const uint32_t r = rank_of(c);
    return LutEntry{r, (3u - r) << ((2 * (k - 1)) & 31)};
}

// ---- memory access: device = read-only global loads, shared-space loads / stores; host build = plain memory ----------
#if defined(__CUDA_ARCH__)
typedef uint32_t saddr_t; // shared-space address (__cvta_generic_to_shared)
K2T_D uint64_t   load8(const uint64_t *p) { return __ldg(p); }
K2T_D uint32_t   bytes_at(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
K2T_D uint64_t   sld64(saddr_t a)
{
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
K2T_D void     sst64(saddr_t a, uint64_t v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
K2T_D LutEntry sld_lut(saddr_t a)
{
    LutEntry e;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.rank), "=r"(e.comp) : "r"(a));
    return e;
}
K2T_D uint32_t opaque(uint32_t x) // keeps a loop-invariant mask in a register instead of a predicate + selects
{
    asm volatile("mov.b32 %0, %0;" : "+r"(x));
    return x;
}
#else
typedef uintptr_t saddr_t;
K2T_D uint64_t    load8(const uint64_t *p) { return *p; }
K2T_D uint32_t    bytes_at(uint32_t a, uint32_t b, uint32_t sel) // PRMT: byte n of the result = byte (sel >> 4n) & 7 of {a, b}
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t       r = 0;
    for (int n = 0; n < 4; ++n)
        r |= (uint32_t)((v >> (8 * ((sel >> (4 * n)) & 7))) & 0xff) << (8 * n);
    return r;
}
K2T_D uint64_t sld64(saddr_t a) { return *reinterpret_cast<const uint64_t *>(a); }
K2T_D void     sst64(saddr_t a, uint64_t v) { *reinterpret_cast<uint64_t *>(a) = v; }
K2T_D LutEntry sld_lut(saddr_t a) { return *reinterpret_cast<const LutEntry *>(a); }
K2T_D uint32_t opaque(uint32_t x) { return x; }
#endif

K2T_D uint64_t min_key(uint64_t a, uint64_t b) { return a < b ? a : b; }

// The bases of a sequence as 4-byte groups cut out of aligned 8-byte words; one word is always in flight.  Touches up to
// 7 bytes before the first base and up to 23 bytes after the last one.
struct BaseStream
{
    const uint64_t *ap;
    uint64_t        x, y; // current word, next word
    uint32_t        sel;
    bool            odd;  // the first group starts in the high half of the first word
    K2T_M void open(const uint8_t *p)
    {
        ap                = reinterpret_cast<const uint64_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)7);
        const uint32_t bo = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 7);
        odd               = bo >= 4;
        sel               = 0x3210u + 0x1111u * (bo & 3);
        x                 = load8(ap);
        y                 = load8(ap + 1);
        ap += 2;
    }
    // the next 8 bases: g0 = bases 0-3, g1 = bases 4-7 (first base in the low byte)
    K2T_M void next8(uint32_t &g0, uint32_t &g1)
    {
        const uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32), y0 = (uint32_t)y, y1 = (uint32_t)(y >> 32);
        const uint32_t A = odd ? x1 : x0, B = odd ? y0 : x1, C = odd ? y1 : y0;
        g0 = bytes_at(A, B, sel);
        g1 = bytes_at(B, C, sel);
        x  = y;
        y  = load8(ap++);
    }
};

// Minimisers of one mate: bases p[0..L), L >= w = k + W - 1.  ring: shared-space address of slot 0 of this thread's W slots,
// slot x at ring + x * stride_bytes; lut: shared-space address of 256 entries of lut_entry(c, k).  Returns the number of
// minimisers; WRITE stores them at out[0..).
//
// SEG: the walk starts inside a longer sequence (segment(), below).  Windows before local k-mer index emit_from are a warm-up
// and emit nothing.  The state machine started afresh holds the right VALUE from its first complete window on (the tracked
// value is always the minimum of the window) but, among equal values, maybe not the reference's POSITION.  The position is
// certain again after a strictly smaller value entered (both walks move to it), or after the tracked value left and the
// minimum of the next window is larger (no equal value was in the window: both walks tracked the one that left, both
// rescan).  *synced: such an event happened at a window in (first complete, emit_from].
template <bool WRITE, bool SEG = false>
K2T_D uint32_t mate(const uint8_t *p, uint32_t L, uint32_t k, uint32_t W, uint64_t seed, uint64_t mask, saddr_t lut, uint64_t *out, saddr_t ring,
                    uint32_t stride_bytes, uint32_t emit_from = 0, bool *synced = nullptr)
{
    const uint32_t mhi = opaque(2 * (k - 1) >= 32 ? ~0u : 0u); // the complement enters the high / the low word
    const uint32_t mlo = ~mhi;
    uint64_t       f = 0, r = 0; // forward / reverse-complement k-mer (kmer_hash.hpp:618-640, minimiser_hash.hpp:91-107)
    auto roll = [&](uint32_t c8) {
        const LutEntry e = sld_lut(lut + c8);
        f                = ((f << 2) | e.rank) & mask;
        r                = (r >> 2) | ((uint64_t)(e.comp & mhi) << 32) | (uint64_t)(e.comp & mlo);
    };

    // phase A: the first k - 1 bases only fill the k-mer registers
    BaseStream bs;
    uint32_t   g0 = 0, g1 = 0;
    {
        bs.open(p);
        for (uint32_t i = 0; i + 1 < k; ++i)
        {
            if ((i & 7) == 0)
                bs.next8(g0, g1);
            else if ((i & 3) == 0)
                g0 = g1;
            roll((g0 & 0xffu) << 3);
            g0 >>= 8;
        }
    }

    // phase B: base k - 1 + j completes k-mer j.  All lanes of a warp are at the same j.
    uint64_t      pre = kBigKey, cur = 0; // prefix minimum (key) of the current block; value of the tracked minimiser
    uint32_t      tag = W - 1;            // W - 1 - offset in the block
    saddr_t       rp  = ring;             // slot of the current offset
    const saddr_t rlast = ring + (saddr_t)(W - 1) * stride_bytes;
    // the tracked minimiser leaves the window when j reaches mq = its index + W.  mq = W - 1 and cur = 0 keep the incomplete
    // windows silent and make the first complete one (j = W - 1) take the rescan branch.
    uint32_t  mq = W - 1;
    uint32_t  bq = 2 * W - 1; // first index of the current block + 2W - 1
    uint32_t  j  = 0;
    uint32_t  n  = 0;
    uint64_t *op = out;
    bool      sync = false;
    auto step = [&](uint32_t c8) {
        roll(c8);
        const uint64_t v   = min_key(f ^ seed, r ^ seed);
        const uint64_t key = (v << kTagBits) | tag;
        pre                = min_key(pre, key);
        sst64(rp, key);
        // window [j-W+1, j]: suffix of the previous block from the next offset on, prefix of this block up to here.  At the
        // last offset the slot read is this block's own first value, which cannot be below the prefix minimum: the prefix wins
        // (so do undefined slots while mq / cur keep the result unused).
        const saddr_t  rn     = rp == rlast ? ring : rp + stride_bytes;
        const uint64_t s      = sld64(rn);
        const bool     take_s = (s | kTagMask) < pre; // strictly smaller value only: ties go to the right block
        const uint64_t wk     = take_s ? s : pre;
        const bool     leave  = j >= mq;              // first window, or the minimiser left (minimiser.hpp:455-461)
        const bool     dec    = v < cur;              // minimiser.hpp:463-468
        const uint32_t mq_w   = bq - ((uint32_t)wk & kTagMask) - (take_s ? W : 0u); // index of the window minimum + W
        const uint64_t was    = cur;
        if (dec)
        {
            cur = v;
            mq  = j + W;
        }
        if (leave)
        {
            cur = wk >> kTagBits;
            mq  = mq_w;
        }
        if (SEG)
            sync |= (dec | (leave & (cur > was))) & (j >= W) & (j <= emit_from);
        if ((leave | dec) && (!SEG || j >= emit_from))
        {
            if (WRITE)
                *op++ = cur;
            else
                ++n;
        }
        ++j;
        rp = rn;
        if (tag == 0)
        {
            // end of the block: the ring becomes the suffix minima of this block (equal values: the later one, smaller tag)
            uint64_t run = key;
            saddr_t  q   = rlast;
#pragma unroll 2
            for (uint32_t c = W - 1; c > 0; --c)
            {
                q -= stride_bytes;
                run = min_key(sld64(q), run);
                sst64(q, run);
            }
            tag = W - 1;
            pre = kBigKey;
            bq += W;
        }
        else
            --tag;
    };
    {
        bs.open(p + (k - 1));
        const uint32_t nk = L - (k - 1); // k-mers = bases left
        uint32_t       i  = 0;
        for (; i + 8 <= nk; i += 8)
        {
            bs.next8(g0, g1);
            step((g0 << 3) & 0x7f8u);
            step((g0 >> 5) & 0x7f8u);
            step((g0 >> 13) & 0x7f8u);
            step((g0 >> 21) & 0x7f8u);
            step((g1 << 3) & 0x7f8u);
            step((g1 >> 5) & 0x7f8u);
            step((g1 >> 13) & 0x7f8u);
            step((g1 >> 21) & 0x7f8u);
        }
        if (i < nk)
        {
            bs.next8(g0, g1);
            for (; i < nk; ++i)
            {
                step((g0 << 3) & 0x7f8u);
                g0 = (g0 >> 8) | (g1 << 24);
                g1 >>= 8;
            }
        }
    }
    if (SEG)
        *synced = sync;
    return WRITE ? (uint32_t)(op - out) : n;
}

// ---- long sequences: segments of kSegWindows windows, one thread each ------------------------------------------------------------
// A sequence of many windows is cut into segments; the thread of segment g walks windows [g * kSegWindows - kSegWarm, end of g)
// and emits from window g * kSegWindows on, at out + g * kSegWindows (every window could emit: the caller's layout has one
// slot per window).  A segment whose warm-up saw no certain state (repeats shorter than the window all along: homopolymers,
// tandem repeats) is flagged in the returned count and its sequence is walked again by one thread from the start.
constexpr uint32_t kSegWindows = 512;
constexpr uint32_t kSegWarm    = 64; // >= 2 * kMaxW: a minimiser lives for at most W windows unless an equal value follows it
constexpr uint32_t kSegFlag    = 0x80000000u;

K2T_D uint32_t segments_of(uint32_t L, uint32_t w) { return L >= w ? (L - w + kSegWindows) / kSegWindows : 0u; }

K2T_D uint32_t segment(const uint8_t *p, uint32_t L, uint32_t seg, uint32_t k, uint32_t w, uint64_t seed, uint64_t mask, saddr_t lut, uint64_t *out,
                       saddr_t ring, uint32_t stride_bytes)
{
    const uint32_t W = w - k + 1, windows = L - w + 1;
    const uint32_t s = seg * kSegWindows, e = windows - s < kSegWindows ? windows : s + kSegWindows;
    const uint32_t warm = s < kSegWarm ? s : kSegWarm, s0 = s - warm;
    bool           synced = false;
    const uint32_t n = mate<true, true>(p + s0, (e - s0) + w - 1, k, W, seed, mask, lut, out + s, ring, stride_bytes, W - 1 + warm, &synced);
    return n | (warm != 0 && !synced ? kSegFlag : 0u);
}

// One read (pair) of a batch: the rules of GanonClassify.cpp:690-700 -- a read shorter than the window is skipped entirely, a
// second mate shorter than the window contributes nothing, hashes(read1) ++ hashes(read2).  KMODE 0: count only; 1: write at
// hash_off[read]; 2: write at hash_off[read] and count (single pass over upper-bound offsets).  Returns the read's total.
template <int KMODE>
K2T_D uint32_t read_pair(uint32_t read, const uint8_t *blk1, const uint32_t *off1, const uint32_t *len1, const uint8_t *blk2, const uint32_t *off2,
                         const uint32_t *len2, uint32_t k, uint32_t w, uint64_t seed, uint64_t mask, saddr_t lut, saddr_t ring, uint32_t stride_bytes,
                         uint32_t *counts, const uint64_t *hash_off, uint64_t *hashes)
{
    constexpr bool WRITE = KMODE != 0;
    const uint32_t W     = w - k + 1;
    uint32_t       total = 0;
    const uint32_t L1    = len1[read];
    if (L1 >= w)
    {
        uint64_t *out = WRITE ? hashes + hash_off[read] : nullptr;
        total         = mate<WRITE>(blk1 + off1[read], L1, k, W, seed, mask, lut, out, ring, stride_bytes);
        if (blk2 != nullptr)
        {
            const uint32_t L2 = len2[read];
            if (L2 >= w)
                total += mate<WRITE>(blk2 + off2[read], L2, k, W, seed, mask, lut, WRITE ? out + total : nullptr, ring, stride_bytes);
        }
    }
    if (KMODE != 1)
        counts[read] = total;
    return total;
}

} // namespace k2t
