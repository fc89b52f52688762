/*
 * synthdb.c -- TEST / BENCHMARK INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Writes the benchmark's synthetic flat database straight to a ganon .ibf file on the CPU, so that the
 * reference arm of bench.py (`--impl reference`) never loads libganon_b200.so: background bits = splitmix64 of the word
 * index (the construction gnb_db_fill_random uses in HBM, restated in ganon_b200/synth.py:random_words) OR-ed with the
 * minimisers of genome g emplaced into bin g (go_minimiser_hash / go_ibf_row of the oracle = IBF.hpp:173-187, 271-286).
 * File layout: save_filter, src/ganon-build/GanonBuild.cpp:251-288 (IBFConfig.hpp:18-40, IBF.hpp:561-571, sdsl
 * int_vector.hpp:2029-2063), the same bytes ganon_b200/formats.py:write_ibf produces.
 *
 *   synthdb OUT.ibf GENOMES.bin bins bin_size h k w genome_len seed target_hashes threads
 *
 * GENOMES.bin: bins * genome_len bytes of ACGT text.  Prints "words xor sum planted" (checksums of the 64-bit words).
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "ganon_oracle.h"

static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

typedef struct
{
    uint64_t widx;
    uint64_t mask;
} plant_t;

static uint64_t  g_bins, g_bin_size, g_bin_words, g_n_words, g_seed, g_chunk_words, g_n_chunks, g_data_off;
static plant_t  *g_plants;
static uint64_t *g_starts; /* [n_chunks + 1] into g_plants (bucketed by chunk) */
static int       g_fd;
static uint64_t  g_next_chunk;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static uint64_t  g_xor, g_sum;
static int       g_failed;

typedef struct
{
    const char *genomes;
    uint64_t    glen, g0, g1, n, cap;
    unsigned    h, k, w;
    plant_t    *out;
} plant_job;

static void *plant_worker(void *arg)
{
    plant_job *J = (plant_job *)arg;
    uint64_t  *mins = (uint64_t *)malloc((J->glen + 1) * 8);
    go_ibf     ibf;
    memset(&ibf, 0, sizeof ibf);
    ibf.bins = g_bins, ibf.technical_bins = g_bin_words * 64, ibf.bin_size = g_bin_size, ibf.bin_words = g_bin_words, ibf.hash_funs = J->h;
    ibf.hash_shift = (uint64_t)__builtin_clzll(g_bin_size);
    const uint64_t kseed = go_adjust_seed(J->k);
    J->cap = (J->g1 - J->g0) * (J->glen / 4 + 16) * J->h + 16;
    J->out = (plant_t *)malloc(J->cap * sizeof(plant_t));
    for (uint64_t g = J->g0; g < J->g1; ++g)
    {
        const size_t nm = go_minimiser_hash(J->genomes + g * J->glen, J->glen, J->k, J->w, kseed, mins);
        for (size_t i = 0; i < nm; ++i)
            for (unsigned fn = 0; fn < J->h; ++fn)
            {
                if (J->n == J->cap)
                {
                    J->cap *= 2;
                    J->out = (plant_t *)realloc(J->out, J->cap * sizeof(plant_t));
                }
                J->out[J->n].widx = go_ibf_row(&ibf, mins[i], fn) * g_bin_words + g / 64;
                J->out[J->n].mask = (uint64_t)1 << (g % 64);
                ++J->n;
            }
    }
    free(mins);
    return NULL;
}

static void *fill_worker(void *arg)
{
    (void)arg;
    uint64_t *buf = (uint64_t *)malloc(g_chunk_words * 8);
    uint64_t  x = 0, s = 0;
    const uint64_t last_word = g_bins / 64, tail_bits = g_bins % 64;
    for (;;)
    {
        pthread_mutex_lock(&g_mu);
        uint64_t c = g_next_chunk++;
        pthread_mutex_unlock(&g_mu);
        if (c >= g_n_chunks)
            break;
        const uint64_t w0 = c * g_chunk_words;
        const uint64_t nw = w0 + g_chunk_words <= g_n_words ? g_chunk_words : g_n_words - w0;
        for (uint64_t i = 0; i < nw; ++i)
            buf[i] = splitmix64(g_seed + (w0 + i) * 8);
        if (tail_bits)
            for (uint64_t i = 0; i < nw; ++i)
            {
                const uint64_t col = (w0 + i) % g_bin_words;
                if (col == last_word)
                    buf[i] &= ((uint64_t)1 << tail_bits) - 1;
                else if (col > last_word)
                    buf[i] = 0;
            }
        for (uint64_t j = g_starts[c]; j < g_starts[c + 1]; ++j)
            buf[g_plants[j].widx - w0] |= g_plants[j].mask;
        for (uint64_t i = 0; i < nw; ++i)
        {
            x ^= buf[i];
            s += buf[i];
        }
        const char *p = (const char *)buf;
        uint64_t    left = nw * 8, off = g_data_off + w0 * 8;
        while (left)
        {
            ssize_t r = pwrite(g_fd, p, left, (off_t)off);
            if (r <= 0)
            {
                g_failed = 1;
                break;
            }
            p += r;
            off += (uint64_t)r;
            left -= (uint64_t)r;
        }
    }
    pthread_mutex_lock(&g_mu);
    g_xor ^= x;
    g_sum += s;
    pthread_mutex_unlock(&g_mu);
    free(buf);
    return NULL;
}

static void w64(FILE *f, uint64_t v) { fwrite(&v, 8, 1, f); }

int main(int argc, char **argv)
{
    if (argc != 12)
    {
        fprintf(stderr, "usage: synthdb OUT.ibf GENOMES.bin bins bin_size h k w genome_len seed target_hashes threads\n");
        return 2;
    }
    const char    *out = argv[1], *gpath = argv[2];
    const uint64_t bins = strtoull(argv[3], 0, 10), bin_size = strtoull(argv[4], 0, 10);
    const unsigned h = (unsigned)atoi(argv[5]), k = (unsigned)atoi(argv[6]), w = (unsigned)atoi(argv[7]);
    const uint64_t glen = strtoull(argv[8], 0, 10), seed = strtoull(argv[9], 0, 10), target_hashes = strtoull(argv[10], 0, 10);
    int            threads = atoi(argv[11]);
    if (threads < 1)
        threads = 1;
    g_bins = bins, g_bin_size = bin_size, g_bin_words = (bins + 63) / 64, g_n_words = bin_size * g_bin_words, g_seed = seed;
    g_chunk_words = (uint64_t)1 << 22; /* 32 MiB */
    g_n_chunks    = (g_n_words + g_chunk_words - 1) / g_chunk_words;

    /* ---- planted bits: genomes split over the threads, each with its own list ---- */
    FILE *gf = fopen(gpath, "rb");
    if (!gf)
    {
        perror(gpath);
        return 1;
    }
    char *genomes = (char *)malloc(bins * glen + 1);
    if (fread(genomes, 1, bins * glen, gf) != bins * glen)
    {
        fprintf(stderr, "synthdb: genomes file too short\n");
        return 1;
    }
    fclose(gf);
    plant_job *jobs = (plant_job *)calloc((size_t)threads, sizeof(plant_job));
    pthread_t *pth = (pthread_t *)malloc((size_t)threads * sizeof(pthread_t));
    for (int t = 0; t < threads; ++t)
    {
        jobs[t].genomes = genomes, jobs[t].glen = glen, jobs[t].h = h, jobs[t].k = k, jobs[t].w = w;
        jobs[t].g0 = bins * (uint64_t)t / (uint64_t)threads, jobs[t].g1 = bins * (uint64_t)(t + 1) / (uint64_t)threads;
        pthread_create(&pth[t], NULL, plant_worker, &jobs[t]);
    }
    uint64_t n_pl = 0;
    for (int t = 0; t < threads; ++t)
    {
        pthread_join(pth[t], NULL);
        n_pl += jobs[t].n;
    }
    free(genomes);
    free(pth);
    /* bucket by chunk (counting sort) */
    g_starts = (uint64_t *)calloc(g_n_chunks + 2, 8);
    for (int t = 0; t < threads; ++t)
        for (uint64_t j = 0; j < jobs[t].n; ++j)
            g_starts[jobs[t].out[j].widx / g_chunk_words + 1]++;
    for (uint64_t c = 0; c < g_n_chunks; ++c)
        g_starts[c + 1] += g_starts[c];
    plant_t  *sorted = (plant_t *)malloc((n_pl + 1) * sizeof(plant_t));
    uint64_t *cur = (uint64_t *)malloc((g_n_chunks + 1) * 8);
    memcpy(cur, g_starts, (g_n_chunks + 1) * 8);
    for (int t = 0; t < threads; ++t)
    {
        for (uint64_t j = 0; j < jobs[t].n; ++j)
            sorted[cur[jobs[t].out[j].widx / g_chunk_words]++] = jobs[t].out[j];
        free(jobs[t].out);
    }
    free(jobs);
    free(cur);
    g_plants = sorted;

    /* ---- header ---- */
    char tmp[4096];
    snprintf(tmp, sizeof tmp, "%s.part", out);
    FILE *f = fopen(tmp, "wb");
    if (!f)
    {
        perror(tmp);
        return 1;
    }
    const int32_t ver[3] = {2, 4, 1};
    fwrite(ver, 4, 3, f);
    w64(f, bins);
    w64(f, target_hashes);
    const uint8_t  h8 = (uint8_t)h, k8 = (uint8_t)k;
    const uint16_t w16 = (uint16_t)w;
    fwrite(&h8, 1, 1, f);
    fwrite(&k8, 1, 1, f);
    fwrite(&w16, 2, 1, f);
    w64(f, bin_size);
    const double fp[3] = {0.05, 0.0, 0.0};
    fwrite(fp, 8, 3, f);
    w64(f, bins); /* hashes_count_std */
    char name[32];
    for (uint64_t b = 0; b < bins; ++b)
    {
        const int n = snprintf(name, sizeof name, "T%llu", (unsigned long long)b);
        w64(f, (uint64_t)n);
        fwrite(name, 1, (size_t)n, f);
        w64(f, target_hashes);
    }
    w64(f, bins); /* bin_map_std */
    for (uint64_t b = 0; b < bins; ++b)
    {
        const int n = snprintf(name, sizeof name, "T%llu", (unsigned long long)b);
        w64(f, b);
        w64(f, (uint64_t)n);
        fwrite(name, 1, (size_t)n, f);
    }
    w64(f, bins);
    w64(f, g_bin_words * 64);
    w64(f, bin_size);
    w64(f, (uint64_t)__builtin_clzll(bin_size));
    w64(f, g_bin_words);
    w64(f, h);
    const uint8_t width = 1;
    const float   growth = 1.5f;
    fwrite(&width, 1, 1, f);
    fwrite(&growth, 4, 1, f);
    w64(f, g_bin_words * 64 * bin_size);
    fflush(f);
    g_data_off = (uint64_t)ftello(f);
    fclose(f);

    /* ---- payload ---- */
    g_fd = open(tmp, O_WRONLY);
    if (g_fd < 0)
    {
        perror(tmp);
        return 1;
    }
    pthread_t *th = (pthread_t *)malloc((size_t)threads * sizeof(pthread_t));
    for (int t = 0; t < threads; ++t)
        pthread_create(&th[t], NULL, fill_worker, NULL);
    for (int t = 0; t < threads; ++t)
        pthread_join(th[t], NULL);
    close(g_fd);
    if (g_failed)
    {
        fprintf(stderr, "synthdb: write failed (disk full?)\n");
        unlink(tmp);
        return 1;
    }
    if (rename(tmp, out) != 0)
    {
        perror(out);
        return 1;
    }
    printf("%llu %llu %llu %llu\n", (unsigned long long)g_n_words, (unsigned long long)g_xor, (unsigned long long)g_sum, (unsigned long long)n_pl);
    return 0;
}
