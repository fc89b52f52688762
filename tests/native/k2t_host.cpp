// Host build of the per-thread core of the K2t kernel (ganon_b200/csrc/k2_thread.cuh) for the CPU test suite: the same
// source the device runs, compiled with g++, so that its logic is checked against the oracle without a GPU.
#include <stdlib.h>
#include <string.h>

#include "../../ganon_b200/csrc/k2_thread.cuh"

extern "C" long k2t_host_minimisers(const uint8_t *seq, uint32_t L, uint32_t k, uint32_t w, uint32_t misalign, uint32_t stride, uint32_t lane,
                                    uint64_t *out, uint64_t cap)
{
    if (k < 1 || k > k2t::kMaxK || w < k || w - k + 1 > k2t::kMaxW || L < w || lane >= stride)
        return -1;
    const uint32_t W = w - k + 1;
    // sequence at an arbitrary byte offset inside an 8-byte aligned buffer, non-sequence bytes on both sides
    uint8_t *raw = (uint8_t *)aligned_alloc(64, ((size_t)L + 128 + 63) / 64 * 64);
    memset(raw, '#', (size_t)L + 128);
    uint8_t *p = raw + 64 + (misalign & 7);
    memcpy(p, seq, L);
    uint64_t *ring = (uint64_t *)malloc((size_t)W * stride * 8);
    memset(ring, 0xA5, (size_t)W * stride * 8); // the kernel's shared memory starts out undefined
    const uint64_t seed = 0x8F3F73B5CF1C9ADEull >> (64 - 2 * k);
    const uint64_t mask = (1ull << (2 * k)) - 1;
    k2t::LutEntry  lut[256];
    for (uint32_t c = 0; c < 256; ++c)
        lut[c] = k2t::lut_entry(c, k);
    const uint32_t n0 = k2t::mate<false>(p, L, k, W, seed, mask, (k2t::saddr_t)lut, nullptr, (k2t::saddr_t)(ring + lane), stride * 8);
    long           rc     = n0;
    if (n0 <= cap)
    {
        memset(ring, 0x5A, (size_t)W * stride * 8);
        const uint32_t n1 = k2t::mate<true>(p, L, k, W, seed, mask, (k2t::saddr_t)lut, out, (k2t::saddr_t)(ring + lane), stride * 8);
        if (n1 != n0)
            rc = -2;
    }
    free(ring);
    free(raw);
    return rc;
}

// A whole batch through k2t::read_pair, the per-thread body of the kernel: blocks with records at arbitrary offsets, optional
// mates, the three kernel modes.  mode 0: counts only; 1: hashes at hash_off; 2: hashes at hash_off (upper bounds) and counts.
extern "C" long k2t_host_batch(const uint8_t *blk1, const uint32_t *off1, const uint32_t *len1, const uint8_t *blk2, const uint32_t *off2, const uint32_t *len2,
                               uint32_t n_reads, uint32_t k, uint32_t w, int mode, uint32_t *counts, const uint64_t *hash_off, uint64_t *hashes)
{
    if (k < 1 || k > k2t::kMaxK || w < k || w - k + 1 > k2t::kMaxW)
        return -1;
    const uint32_t W = w - k + 1;
    const uint32_t T = k2t::kThreads;
    uint64_t      *ring = (uint64_t *)malloc((size_t)W * T * 8);
    memset(ring, 0xA5, (size_t)W * T * 8);
    const uint64_t seed = 0x8F3F73B5CF1C9ADEull >> (64 - 2 * k);
    const uint64_t mask = (1ull << (2 * k)) - 1;
    k2t::LutEntry  lut[256];
    for (uint32_t c = 0; c < 256; ++c)
        lut[c] = k2t::lut_entry(c, k);
    unsigned long long sum = 0;
    for (uint32_t read = 0; read < n_reads; ++read)
    {
        const k2t::saddr_t ring_s = (k2t::saddr_t)(ring + read % T); // the slot column of thread read % T
        uint32_t           total;
        if (mode == 0)
            total = k2t::read_pair<0>(read, blk1, off1, len1, blk2, off2, len2, k, w, seed, mask, (k2t::saddr_t)lut, ring_s, T * 8, counts, hash_off, hashes);
        else if (mode == 1)
            total = k2t::read_pair<1>(read, blk1, off1, len1, blk2, off2, len2, k, w, seed, mask, (k2t::saddr_t)lut, ring_s, T * 8, counts, hash_off, hashes);
        else
            total = k2t::read_pair<2>(read, blk1, off1, len1, blk2, off2, len2, k, w, seed, mask, (k2t::saddr_t)lut, ring_s, T * 8, counts, hash_off, hashes);
        sum += total;
    }
    free(ring);
    return (long)sum;
}

// A long sequence through k2t::segment, one call per segment as the segment kernel's threads make them: out has one slot per
// window (segment g writes from slot g * kSegWindows), seg_cnt[g] = count | kSegFlag.  Returns the number of segments.
extern "C" long k2t_host_segments(const uint8_t *seq, uint32_t L, uint32_t k, uint32_t w, uint32_t misalign, uint32_t stride, uint32_t lane, uint64_t *out,
                                  uint32_t *seg_cnt)
{
    if (k < 1 || k > k2t::kMaxK || w < k || w - k + 1 > k2t::kMaxW || L < w || lane >= stride)
        return -1;
    const uint32_t W   = w - k + 1;
    uint8_t       *raw = (uint8_t *)aligned_alloc(64, ((size_t)L + 128 + 63) / 64 * 64);
    memset(raw, '#', (size_t)L + 128);
    uint8_t *p = raw + 64 + (misalign & 7);
    memcpy(p, seq, L);
    uint64_t      *ring = (uint64_t *)malloc((size_t)W * stride * 8);
    const uint64_t seed = 0x8F3F73B5CF1C9ADEull >> (64 - 2 * k);
    const uint64_t mask = (1ull << (2 * k)) - 1;
    k2t::LutEntry  lut[256];
    for (uint32_t c = 0; c < 256; ++c)
        lut[c] = k2t::lut_entry(c, k);
    const uint32_t n_seg = k2t::segments_of(L, w);
    for (uint32_t g = 0; g < n_seg; ++g)
    {
        memset(ring, 0xA5 ^ g, (size_t)W * stride * 8);
        seg_cnt[g] = k2t::segment(p, L, g, k, w, seed, mask, (k2t::saddr_t)lut, out, (k2t::saddr_t)(ring + lane), stride * 8);
    }
    free(ring);
    free(raw);
    return n_seg;
}
