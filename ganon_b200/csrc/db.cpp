// Database files -> HBM.  Hand-written readers for the cereal-binary layouts the reference writes:
//   .ibf  : save_filter GanonBuild.cpp:251-288 / load_filter GC.cpp:949-986 (IBFConfig.hpp:18-40, IBF.hpp:561-571,
//           sdsl int_vector.hpp:2029-2035)
//   .hibf : raptor 3.0.1 index, load_filter GC.cpp:875-938 (HIBF.hpp:163-169, 293-298; seqan3::shape =
//           dynamic_bitset.hpp:1963-1972)
// The bitvector payload is streamed file -> two pinned staging buffers -> HBM; with n_shards > 1 only the shard's
// bin-word columns are copied (cudaMemcpy2DAsync picks the column slice out of each staged run of rows).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <functional>
#include <memory>
#include <sys/stat.h>
#include <unistd.h>
#include <unordered_map>

#include "db.h"

namespace gnb
{

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int  fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

namespace
{

struct FileReader
{
    int      fd  = -1;
    uint64_t pos = 0, size = 0;
    bool     ok  = true;
    bool     open(const char *path)
    {
        fd = ::open(path, O_RDONLY);
        if (fd < 0)
            return false;
        struct stat st;
        if (fstat(fd, &st) != 0)
            return false;
        size = (uint64_t)st.st_size;
        return true;
    }
    ~FileReader()
    {
        if (fd >= 0)
            ::close(fd);
    }
    bool read(void *dst, uint64_t n)
    {
        uint8_t *d = (uint8_t *)dst;
        while (n)
        {
            ssize_t r = ::pread(fd, d, n > (1u << 30) ? (1u << 30) : n, (off_t)pos);
            if (r <= 0)
            {
                ok = false;
                return false;
            }
            d += r;
            pos += (uint64_t)r;
            n -= (uint64_t)r;
        }
        return true;
    }
    template <typename T>
    T get()
    {
        T v{};
        read(&v, sizeof(T));
        return v;
    }
    std::string str()
    {
        uint64_t n = get<uint64_t>();
        if (!ok || n > size - std::min(size, pos))
        {
            ok = false;
            return {};
        }
        std::string s(n, '\0');
        if (n)
            read(&s[0], n);
        return s;
    }
};

// IBF body as serialised by IBF.hpp:561-571: six u64 + sdsl bit_vector (u8 width, f32 growth, u64 bits, words)
} // namespace
bool plan_pages(const IbfHost &t, uint64_t budget, std::vector<IbfPage> &pages, size_t &n_resident);
int  alloc_page_storage(IbfHost &t, size_t n_resident);
namespace
{
int read_ibf_body(FileReader &f, IbfHost &ibf, int shard, int n_shards, cudaStream_t st, void *pinned[2], cudaEvent_t ev[2], size_t pinned_bytes,
                  uint64_t hbm_budget = 0)
{
    ibf.bins           = f.get<uint64_t>();
    ibf.technical_bins = f.get<uint64_t>();
    ibf.bin_size       = f.get<uint64_t>();
    ibf.hash_shift     = f.get<uint64_t>();
    ibf.bin_words      = f.get<uint64_t>();
    ibf.hash_funs      = f.get<uint64_t>();
    const uint8_t width = f.get<uint8_t>();
    (void)f.get<float>();
    const uint64_t n_bits = f.get<uint64_t>();
    if (!f.ok)
        return fail(GNB_ERR_IO, "truncated IBF header");
    // (untrusted file: the products are checked for overflow before they are compared with anything)
    uint64_t tech_bits = 0, want_bits = 0;
    if (width != 1 || ibf.bin_words == 0 || ibf.bin_size == 0 || __builtin_mul_overflow(ibf.bin_words, (uint64_t)64, &tech_bits) || ibf.technical_bins != tech_bits ||
        ibf.bins > ibf.technical_bins || __builtin_mul_overflow(ibf.technical_bins, ibf.bin_size, &want_bits) || n_bits != want_bits || ibf.hash_funs < 1 ||
        ibf.hash_funs > 5 || ibf.hash_shift != (uint64_t)__builtin_clzll(ibf.bin_size))
        return fail(GNB_ERR_FORMAT, "inconsistent interleaved_bloom_filter header");
    const uint64_t n_words = n_bits / 64 + (n_bits % 64 ? 1 : 0);
    if (f.pos > f.size || n_words > (f.size - f.pos) / 8)
        return fail(GNB_ERR_IO, "truncated IBF payload");
    ibf.w0 = ibf.bin_words * (uint64_t)shard / (uint64_t)n_shards;
    ibf.w1 = ibf.bin_words * (uint64_t)(shard + 1) / (uint64_t)n_shards;
    if (ibf.w1 <= ibf.w0)
        return fail(GNB_ERR_ARG, "more shards than bin-words");
    if (ibf.row_words() >= (1ull << 31))
        return fail(GNB_ERR_LIMIT, "row too wide");
    const uint64_t row_bytes  = ibf.bin_words * 8;
    uint64_t       rows_per   = pinned_bytes / row_bytes;
    if (rows_per == 0)
        return fail(GNB_ERR_LIMIT, "row larger than the staging buffer");
    int      cur = 0;
    uint64_t row = 0;
    if (hbm_budget && n_shards == 1 && ibf.device_bytes() > hbm_budget)
    {
        // host-resident tier: one pass over the file, every row cut into its column pages -- resident pages go to HBM,
        // the others to page-locked host memory
        size_t n_res = 0;
        if (!plan_pages(ibf, hbm_budget, ibf.pages, n_res))
            return fail(GNB_ERR_LIMIT, "the HBM budget does not hold two one-word pages of this filter");
        GNB_TRY(alloc_page_storage(ibf, n_res));
        while (row < ibf.bin_size)
        {
            const uint64_t rows = std::min(rows_per, ibf.bin_size - row);
            GNB_CUDA(cudaEventSynchronize(ev[cur]));
            if (!f.read(pinned[cur], rows * row_bytes))
                return fail(GNB_ERR_IO, "short read in IBF payload");
            for (auto &p : ibf.pages)
            {
                const uint64_t pw = p.w1 - p.w0;
                const uint8_t *src = (const uint8_t *)pinned[cur] + p.w0 * 8;
                if (p.d_data)
                    GNB_CUDA(cudaMemcpy2DAsync(p.d_data + row * pw, pw * 8, src, row_bytes, pw * 8, rows, cudaMemcpyHostToDevice, st));
                else
                    for (uint64_t r = 0; r < rows; ++r)
                        memcpy(p.h_data + (row + r) * pw, src + r * row_bytes, pw * 8);
            }
            GNB_CUDA(cudaEventRecord(ev[cur], st));
            cur ^= 1;
            row += rows;
        }
        return GNB_OK;
    }
    GNB_CUDA(cudaMalloc((void **)&ibf.d_data, ibf.device_bytes()));
    while (row < ibf.bin_size)
    {
        const uint64_t rows = std::min(rows_per, ibf.bin_size - row);
        GNB_CUDA(cudaEventSynchronize(ev[cur]));
        if (!f.read(pinned[cur], rows * row_bytes))
            return fail(GNB_ERR_IO, "short read in IBF payload");
        if (n_shards == 1)
            GNB_CUDA(cudaMemcpyAsync(ibf.d_data + row * ibf.bin_words, pinned[cur], rows * row_bytes, cudaMemcpyHostToDevice, st));
        else
            GNB_CUDA(cudaMemcpy2DAsync(ibf.d_data + row * ibf.row_words(), ibf.row_words() * 8, (const uint8_t *)pinned[cur] + ibf.w0 * 8,
                                       row_bytes, ibf.row_words() * 8, rows, cudaMemcpyHostToDevice, st));
        GNB_CUDA(cudaEventRecord(ev[cur], st));
        cur ^= 1;
        row += rows;
    }
    return GNB_OK;
}

} // namespace

bool plan_pages(const IbfHost &t, uint64_t budget, std::vector<IbfPage> &pages, size_t &n_resident)
{
    pages.clear();
    n_resident           = 0;
    const uint64_t words = t.row_words(), col_bytes = t.bin_size * 8; // bytes of one bin-word column
    // a page takes at most 1/8 of the budget (two of them are staging buffers), whole 64-word chunks where the row allows
    uint64_t pw = budget / 8 / col_bytes;
    if (pw >= 64)
        pw = pw / 64 * 64;
    if (pw == 0)
        pw = budget / 2 / col_bytes; // tiny budgets (tests): at least the two staging buffers must fit
    if (pw == 0)
        return false;
    pw = std::min(pw, words);
    for (uint64_t w = 0; w < words; w += pw)
    {
        IbfPage p;
        p.w0 = t.w0 + w;
        p.w1 = t.w0 + std::min(words, w + pw);
        pages.push_back(p);
    }
    const uint64_t fit = budget / (pw * col_bytes);
    n_resident         = fit > 2 ? (size_t)std::min<uint64_t>(fit - 2, pages.size()) : 0;
    return true;
}

int alloc_page_storage(IbfHost &t, size_t n_resident)
{
    t.stage_bytes = 0;
    for (size_t i = n_resident; i < t.pages.size(); ++i)
        t.stage_bytes = std::max(t.stage_bytes, t.page_bytes(t.pages[i]));
    for (size_t i = 0; i < t.pages.size(); ++i)
    {
        IbfPage &p = t.pages[i];
        if (i < n_resident)
            GNB_CUDA(cudaMalloc((void **)&p.d_data, t.page_bytes(p)));
        else
            GNB_CUDA(cudaMallocHost((void **)&p.h_data, t.page_bytes(p)));
    }
    if (t.stage_bytes)
    {
        GNB_CUDA(cudaStreamCreateWithFlags(&t.copy_st, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b)
        {
            GNB_CUDA(cudaMalloc((void **)&t.d_stage[b], t.stage_bytes));
            GNB_CUDA(cudaEventCreateWithFlags(&t.ev_ready[b], cudaEventDisableTiming));
            GNB_CUDA(cudaEventCreateWithFlags(&t.ev_free[b], cudaEventDisableTiming));
        }
    }
    return GNB_OK;
}

void free_page_storage(IbfHost &t)
{
    for (auto &p : t.pages)
    {
        if (p.d_data)
            cudaFree(p.d_data);
        if (p.h_data)
            cudaFreeHost(p.h_data);
    }
    t.pages.clear();
    for (int b = 0; b < 2; ++b)
    {
        if (t.d_stage[b])
            cudaFree(t.d_stage[b]);
        if (t.ev_ready[b])
            cudaEventDestroy(t.ev_ready[b]);
        if (t.ev_free[b])
            cudaEventDestroy(t.ev_free[b]);
        t.d_stage[b]  = nullptr;
        t.ev_ready[b] = t.ev_free[b] = nullptr;
    }
    if (t.copy_st)
        cudaStreamDestroy(t.copy_st);
    t.copy_st = nullptr;
}

namespace
{
void replace_all(std::string &s, const std::string &from, const std::string &to)
{
    size_t p = 0;
    while ((p = s.find(from, p)) != std::string::npos)
    {
        s.replace(p, from.size(), to);
        p += to.size();
    }
}

} // namespace
} // namespace gnb

using namespace gnb;

// filter.map (GC.cpp:1021-1025) and target_fpr (flat: GC.cpp:969-982 with false_positive 940-947; HIBF: GC.cpp:932)
void gnb_db::derive_targets()
{
    target_names.clear();
    target_bins.clear();
    target_fpr.clear();
    std::unordered_map<std::string, size_t> idx;
    for (auto const &[binno, name] : bin_map)
    {
        auto it = idx.find(name);
        if (it == idx.end())
        {
            it = idx.emplace(name, target_names.size()).first;
            target_names.push_back(name);
            target_bins.emplace_back();
        }
        target_bins[it->second].push_back(binno);
    }
    target_fpr.assign(target_names.size(), is_hibf ? max_fp : 0.0);
    if (!is_hibf)
    {
        const uint64_t bin_size_bits = ibfs[0].bin_size;
        const uint8_t  hash_functions = (uint8_t)ibfs[0].hash_funs;
        for (auto const &[target, count] : hashes_count)
        {
            auto it = idx.find(target);
            if (it == idx.end() || count == 0 || max_hashes_bin == 0)
                continue;
            uint64_t n_bins_target = (uint64_t)std::ceil(count / static_cast<double>(max_hashes_bin));
            uint64_t n_hashes_bin  = (uint64_t)std::ceil(count / static_cast<double>(n_bins_target));
            double   fp = std::pow(1 - std::exp(-hash_functions / (bin_size_bits / static_cast<double>(n_hashes_bin))), hash_functions);
            target_fpr[it->second] = 1.0 - std::pow(1.0 - fp, n_bins_target);
        }
    }
}

extern "C" const char *gnb_last_error(void) { return gnb::g_err.c_str(); }
extern "C" int         gnb_abi_version(void) { return GNB_ABI_VERSION; }

extern "C" int gnb_device_count(int *n)
{
    if (!n)
        return fail(GNB_ERR_ARG, "null argument");
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess)
    {
        *n = 0;
        return fail(GNB_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    return GNB_OK;
}

static int open_impl(const char *path, int is_hibf, int device, int shard, int n_shards, uint64_t hbm_budget, gnb_db **out);

extern "C" int gnb_db_open(const char *path, int is_hibf, int device, int shard, int n_shards, gnb_db **out)
{
    // GANON_B200_HBM_BUDGET_GB: flat filters larger than this are loaded in the paged form (host-resident tier)
    uint64_t budget = 0;
    if (const char *e = getenv("GANON_B200_HBM_BUDGET_GB"))
        budget = (uint64_t)(atof(e) * (double)(1ull << 30));
    return open_impl(path, is_hibf, device, shard, n_shards, n_shards == 1 && !is_hibf ? budget : 0, out);
}

extern "C" int gnb_db_open_paged(const char *path, int device, uint64_t hbm_budget_bytes, gnb_db **out)
{
    return open_impl(path, 0, device, 0, 1, hbm_budget_bytes, out);
}

static int open_impl(const char *path, int is_hibf, int device, int shard, int n_shards, uint64_t hbm_budget, gnb_db **out)
{
    if (!path || !out || n_shards < 1 || shard < 0 || shard >= n_shards)
        return fail(GNB_ERR_ARG, "gnb_db_open: bad arguments");
    if (is_hibf && n_shards != 1)
        return fail(GNB_ERR_ARG, "gnb_db_open: bin-block sharding is implemented for flat .ibf only");
    *out = nullptr;
    FileReader f;
    if (!f.open(path))
        return fail(GNB_ERR_IO, std::string("file not found: ") + path);
    if (f.size == 0)
        return fail(GNB_ERR_IO, std::string("file is empty: ") + path);
    GNB_CUDA(cudaSetDevice(device));
    std::unique_ptr<gnb_db> db(new gnb_db);
    db->is_hibf = is_hibf != 0;
    db->device  = device;

    const size_t pinned_bytes = 64u << 20;
    void        *pinned[2]    = {nullptr, nullptr};
    cudaEvent_t  ev[2]        = {nullptr, nullptr};
    cudaStream_t st           = nullptr;
    bool         cleaned      = false;
    auto cleanup = [&]() {
        if (cleaned)
            return;
        cleaned = true;
        if (st)
            cudaStreamSynchronize(st);
        for (int i = 0; i < 2; ++i)
        {
            if (pinned[i])
                cudaFreeHost(pinned[i]);
            if (ev[i])
                cudaEventDestroy(ev[i]);
        }
        if (st)
            cudaStreamDestroy(st);
    };
    // every exit path below releases the staging resources and, through gnb_db_free, whatever the handle already holds
    struct Guard
    {
        std::function<void()> fn;
        std::unique_ptr<gnb_db> &db;
        ~Guard()
        {
            fn();
            if (db)
                gnb_db_free(db.release());
        }
    } guard{cleanup, db};
    GNB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i)
    {
        GNB_CUDA(cudaMallocHost(&pinned[i], pinned_bytes));
        GNB_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    }
    int rc = GNB_OK;
    if (!is_hibf)
    {
        for (int i = 0; i < 3; ++i)
            db->version[i] = f.get<int32_t>();
        // IBFConfig (IBFConfig.hpp:18-40)
        const uint64_t n_bins       = f.get<uint64_t>();
        db->max_hashes_bin          = f.get<uint64_t>();
        const uint8_t  hash_funs    = f.get<uint8_t>();
        db->kmer_size               = f.get<uint8_t>();
        db->window_size             = f.get<uint16_t>();
        const uint64_t bin_size_bits = f.get<uint64_t>();
        db->max_fp                  = f.get<double>();
        db->true_max_fp             = f.get<double>();
        db->true_avg_fp             = f.get<double>();
        uint64_t n = f.get<uint64_t>();
        if (!f.ok || n > f.size)
            rc = fail(GNB_ERR_FORMAT, std::string("not a ganon .ibf file: ") + path);
        for (uint64_t i = 0; rc == GNB_OK && i < n; ++i)
        {
            std::string t = f.str();
            uint64_t    c = f.get<uint64_t>();
            db->hashes_count.emplace_back(std::move(t), c);
            if (!f.ok)
                rc = fail(GNB_ERR_FORMAT, "truncated hashes_count");
        }
        uint64_t m = rc == GNB_OK ? f.get<uint64_t>() : 0;
        if (rc == GNB_OK && (!f.ok || m > f.size))
            rc = fail(GNB_ERR_FORMAT, "truncated bin_map");
        for (uint64_t i = 0; rc == GNB_OK && i < m; ++i)
        {
            uint64_t    b = f.get<uint64_t>();
            std::string t = f.str();
            db->bin_map.emplace_back(b, std::move(t));
            if (!f.ok)
                rc = fail(GNB_ERR_FORMAT, "truncated bin_map");
        }
        if (rc == GNB_OK)
        {
            db->ibfs.resize(1);
            rc = read_ibf_body(f, db->ibfs[0], shard, n_shards, st, pinned, ev, pinned_bytes, hbm_budget);
            if (rc == GNB_OK && (db->ibfs[0].bin_size != bin_size_bits || db->ibfs[0].hash_funs != hash_funs || db->ibfs[0].bins != n_bins))
                rc = fail(GNB_ERR_FORMAT, "IBFConfig does not match the filter");
            for (auto const &bm : db->bin_map)
                if (rc == GNB_OK && bm.first >= db->ibfs[0].technical_bins)
                    rc = fail(GNB_ERR_FORMAT, "bin_map entry out of range");
        }
    }
    else
    {
        (void)f.get<uint32_t>(); // index version
        db->window_size            = (uint32_t)f.get<uint64_t>();
        const uint64_t shape_size  = f.get<uint64_t>();
        const uint64_t shape_bits  = f.get<uint64_t>();
        (void)shape_size;
        db->kmer_size = (uint32_t)__builtin_popcountll(shape_bits);
        (void)f.get<uint8_t>(); // parts
        (void)f.get<uint8_t>(); // compressed
        uint64_t n_ub = f.get<uint64_t>();
        if (!f.ok || n_ub > f.size)
            rc = fail(GNB_ERR_FORMAT, std::string("not a raptor .hibf file: ") + path);
        for (uint64_t u = 0; rc == GNB_OK && u < n_ub; ++u)
        {
            uint64_t nf = f.get<uint64_t>();
            if (!f.ok || nf > f.size)
            {
                rc = fail(GNB_ERR_FORMAT, "truncated bin_path");
                break;
            }
            for (uint64_t j = 0; j < nf; ++j)
            {
                std::string p = f.str();
                // target name from the path (GC.cpp:908-925)
                size_t      sl = p.find_last_of('/');
                std::string t  = sl == std::string::npos ? p : p.substr(sl + 1);
                size_t      mi = t.find(".minimiser");
                if (mi != std::string::npos)
                    t = t.substr(0, mi);
                replace_all(t, "|||", ".");
                replace_all(t, "---", " ");
                db->bin_map.emplace_back(u, std::move(t));
            }
        }
        db->max_fp = f.get<double>();
        (void)f.get<uint8_t>(); // is_hibf
        uint64_t n_ibf = f.get<uint64_t>();
        if (rc == GNB_OK && (!f.ok || n_ibf == 0 || n_ibf > f.size))
            rc = fail(GNB_ERR_FORMAT, "truncated hibf");
        if (rc == GNB_OK)
            db->ibfs.resize(n_ibf);
        for (uint64_t i = 0; rc == GNB_OK && i < n_ibf; ++i)
            rc = read_ibf_body(f, db->ibfs[i], 0, 1, st, pinned, ev, pinned_bytes);
        auto read_vv = [&](std::vector<std::vector<int64_t>> &vv) {
            uint64_t n = f.get<uint64_t>();
            if (!f.ok || n > (f.size - std::min(f.size, f.pos)) / 8)
            {
                rc = fail(GNB_ERR_FORMAT, "truncated hibf tables");
                return;
            }
            vv.resize(n);
            for (auto &v : vv)
            {
                uint64_t m = f.get<uint64_t>();
                if (!f.ok || m > (f.size - std::min(f.size, f.pos)) / 8)
                {
                    rc = fail(GNB_ERR_FORMAT, "truncated hibf tables");
                    return;
                }
                v.resize(m);
                if (m)
                    f.read(v.data(), m * 8);
            }
        };
        if (rc == GNB_OK)
            read_vv(db->next_ibf_id);
        if (rc == GNB_OK)
        {
            uint64_t n = f.get<uint64_t>(); // user_bin_filenames
            for (uint64_t i = 0; f.ok && i < n; ++i)
                (void)f.str();
        }
        if (rc == GNB_OK)
            read_vv(db->bin_to_user);
        if (rc == GNB_OK && (!f.ok || db->next_ibf_id.size() != db->ibfs.size() || db->bin_to_user.size() != db->ibfs.size()))
            rc = fail(GNB_ERR_FORMAT, "inconsistent hibf tables");
        db->n_user_bins = n_ub;
        if (rc == GNB_OK)
            for (size_t i = 0; i < db->ibfs.size(); ++i)
            {
                if (db->next_ibf_id[i].size() < db->ibfs[i].bins || db->bin_to_user[i].size() < db->ibfs[i].bins)
                    rc = fail(GNB_ERR_FORMAT, "hibf tables shorter than the bin count");
                for (uint64_t b = 0; rc == GNB_OK && b < db->ibfs[i].bins; ++b)
                {
                    const int64_t fi = db->bin_to_user[i][b], nx = db->next_ibf_id[i][b];
                    if (fi >= (int64_t)n_ub || (fi < 0 && (nx < 0 || nx >= (int64_t)db->ibfs.size())))
                        rc = fail(GNB_ERR_FORMAT, "hibf table entry out of range");
                }
            }
    }
    cleanup();
    if (rc != GNB_OK)
    {
        std::string keep = g_err;
        gnb_db_free(db.release());
        g_err = keep;
        return rc;
    }
    db->derive_targets();
    *out = db.release();
    return GNB_OK;
}

extern "C" void gnb_db_free(gnb_db *db)
{
    if (!db)
        return;
    cudaSetDevice(db->device);
    for (auto &i : db->ibfs)
    {
        if (i.d_data)
            cudaFree(i.d_data);
        free_page_storage(i);
    }
    delete db;
}

extern "C" int gnb_db_info(const gnb_db *db, gnb_db_info_t *info)
{
    if (!db || !info)
        return fail(GNB_ERR_ARG, "null argument");
    const IbfHost &t = db->ibfs[0];
    info->is_hibf        = db->is_hibf;
    info->kmer_size      = db->kmer_size;
    info->window_size    = db->window_size;
    info->hash_functions = (uint32_t)t.hash_funs;
    info->bins           = t.bins;
    info->technical_bins = t.technical_bins;
    info->bin_size_bits  = t.bin_size;
    info->bin_words      = t.bin_words;
    info->shard_word_begin = t.w0;
    info->shard_word_end   = t.w1;
    info->max_hashes_bin = db->max_hashes_bin;
    info->max_fp         = db->max_fp;
    info->n_targets      = db->target_names.size();
    info->n_ibfs         = db->ibfs.size();
    info->device_bytes   = 0;
    for (auto const &i : db->ibfs)
        info->device_bytes += i.device_bytes();
    info->device = db->device;
    info->n_pages = info->n_resident_pages = 0;
    info->host_bytes = 0;
    for (auto const &i : db->ibfs)
    {
        info->n_pages += i.pages.size();
        info->host_bytes += i.host_bytes();
        for (auto const &p : i.pages)
            info->n_resident_pages += p.d_data ? 1 : 0;
    }
    return GNB_OK;
}

// Host-resident tier: turn a filter that is whole in HBM into its paged form for `hbm_budget_bytes` (tests and the
// benchmark build their synthetic databases in HBM first; files that never fit are loaded paged by gnb_db_open_paged).
extern "C" int gnb_db_page_out(gnb_db *db, uint64_t hbm_budget_bytes)
{
    if (!db || db->is_hibf || db->ibfs.size() != 1)
        return fail(GNB_ERR_ARG, "gnb_db_page_out: a flat IBF is needed");
    IbfHost &t = db->ibfs[0];
    if (t.paged() || !t.d_data)
        return fail(GNB_ERR_ARG, "gnb_db_page_out: the filter is already paged");
    GNB_CUDA(cudaSetDevice(db->device));
    std::vector<IbfPage> pages;
    size_t               n_res = 0;
    if (t.bin_size * t.row_words() * 8 <= hbm_budget_bytes)
        return GNB_OK; // fits: nothing to do
    if (!plan_pages(t, hbm_budget_bytes, pages, n_res))
        return fail(GNB_ERR_LIMIT, "gnb_db_page_out: the budget does not hold two one-word pages of this filter");
    t.pages = pages;
    int rc  = alloc_page_storage(t, n_res);
    if (rc != GNB_OK)
    {
        free_page_storage(t);
        return rc;
    }
    const uint64_t pitch = t.row_words() * 8;
    for (auto &p : t.pages)
    {
        const uint64_t width = (p.w1 - p.w0) * 8;
        const void    *src   = reinterpret_cast<const uint8_t *>(t.d_data) + (p.w0 - t.w0) * 8;
        cudaError_t    e     = cudaMemcpy2D(p.d_data ? (void *)p.d_data : (void *)p.h_data, width, src, pitch, width, t.bin_size,
                                            p.d_data ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost);
        if (e != cudaSuccess)
        {
            free_page_storage(t);
            return fail(GNB_ERR_CUDA, std::string("gnb_db_page_out: ") + cudaGetErrorString(e));
        }
    }
    cudaFree(t.d_data);
    t.d_data = nullptr;
    return GNB_OK;
}

extern "C" int gnb_db_target(const gnb_db *db, uint64_t i, const char **name, double *fpr, uint64_t *n_bins)
{
    if (!db || i >= db->target_names.size())
        return fail(GNB_ERR_ARG, "target index out of range");
    if (name)
        *name = db->target_names[i].c_str();
    if (fpr)
        *fpr = db->target_fpr[i];
    if (n_bins)
        *n_bins = db->target_bins[i].size();
    return GNB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// build side
// ------------------------------------------------------------------------------------------------------------------
extern "C" int gnb_db_create(uint64_t bins, uint64_t bin_size_bits, uint32_t hash_functions, uint32_t kmer_size, uint32_t window_size,
                             int device, gnb_db **out)
{
    return gnb_db_create_sharded(bins, bin_size_bits, hash_functions, kmer_size, window_size, device, 0, 1, out);
}

extern "C" int gnb_db_create_sharded(uint64_t bins, uint64_t bin_size_bits, uint32_t hash_functions, uint32_t kmer_size, uint32_t window_size,
                                     int device, int shard, int n_shards, gnb_db **out)
{
    if (!out || n_shards < 1 || shard < 0 || shard >= n_shards || bins == 0 || bin_size_bits == 0 || hash_functions < 1 || hash_functions > 5 || kmer_size < 1 || kmer_size > 32 ||
        window_size < kmer_size)
        return fail(GNB_ERR_ARG, "gnb_db_create: bad arguments"); // IBF.hpp:227-236
    GNB_CUDA(cudaSetDevice(device));
    std::unique_ptr<gnb_db> db(new gnb_db);
    db->device      = device;
    db->kmer_size   = kmer_size;
    db->window_size = window_size;
    db->max_fp      = 0.05;
    db->ibfs.resize(1);
    IbfHost &t       = db->ibfs[0];
    t.bins           = bins;
    t.bin_words      = (bins + 63) >> 6;
    t.technical_bins = t.bin_words << 6;
    t.bin_size       = bin_size_bits;
    t.hash_shift     = (uint64_t)__builtin_clzll(bin_size_bits);
    t.hash_funs      = hash_functions;
    t.w0             = t.bin_words * (uint64_t)shard / (uint64_t)n_shards;
    t.w1             = t.bin_words * (uint64_t)(shard + 1) / (uint64_t)n_shards;
    if (t.w1 <= t.w0)
        return fail(GNB_ERR_ARG, "more shards than bin-words");
    GNB_CUDA(cudaMalloc((void **)&t.d_data, t.device_bytes()));
    GNB_CUDA(cudaMemset(t.d_data, 0, t.device_bytes()));
    // default map: one target per bin, "T<bin>"
    db->max_hashes_bin = 1;
    for (uint64_t b = 0; b < bins; ++b)
        db->bin_map.emplace_back(b, "T" + std::to_string(b));
    for (uint64_t b = 0; b < bins; ++b)
        db->hashes_count.emplace_back("T" + std::to_string(b), 1);
    db->derive_targets();
    *out = db.release();
    return GNB_OK;
}

extern "C" int gnb_db_fill_random(gnb_db *db, uint64_t seed, int and_terms)
{
    if (!db || and_terms < 0 || and_terms > 16)
        return fail(GNB_ERR_ARG, "gnb_db_fill_random: bad arguments");
    if (db->ibfs[0].paged())
        return fail(GNB_ERR_ARG, "not available on a paged filter (host-resident tier): the bitvector is not whole in HBM");
    GNB_CUDA(cudaSetDevice(db->device));
    for (size_t i = 0; i < db->ibfs.size(); ++i)
    { // sub-IBF i of an HIBF draws from seed + i
        IbfHost &t = db->ibfs[i];
        launch_fill_random(t.d_data, t.bin_size, (uint32_t)t.row_words(), (uint32_t)t.w0, (uint32_t)t.bin_words, t.bins, seed + i, and_terms, 0);
    }
    GNB_CUDA(cudaGetLastError());
    GNB_CUDA(cudaDeviceSynchronize());
    return GNB_OK;
}

extern "C" int gnb_db_emplace(gnb_db *db, const uint64_t *hashes, const uint32_t *bins, uint64_t n)
{
    if (!db || db->is_hibf)
        return fail(GNB_ERR_ARG, "gnb_db_emplace: bad arguments");
    return gnb_db_emplace_ibf(db, 0, hashes, bins, n);
}

extern "C" int gnb_db_emplace_ibf(gnb_db *db, uint64_t ibf_index, const uint64_t *hashes, const uint32_t *bins, uint64_t n)
{
    if (!db || ibf_index >= db->ibfs.size() || (n && (!hashes || !bins)))
        return fail(GNB_ERR_ARG, "gnb_db_emplace: bad arguments");
    if (db->ibfs[0].paged())
        return fail(GNB_ERR_ARG, "not available on a paged filter (host-resident tier): the bitvector is not whole in HBM");
    if (n == 0)
        return GNB_OK;
    GNB_CUDA(cudaSetDevice(db->device));
    IbfHost &t = db->ibfs[ibf_index];
    for (uint64_t i = 0; i < n; ++i)
        if (bins[i] >= t.bins)
            return fail(GNB_ERR_ARG, "gnb_db_emplace: bin out of range"); // IBF.hpp:274
    uint64_t *d_h = nullptr;
    uint32_t *d_b = nullptr;
    GNB_CUDA(cudaMalloc((void **)&d_h, n * 8));
    GNB_CUDA(cudaMalloc((void **)&d_b, n * 4));
    GNB_CUDA(cudaMemcpy(d_h, hashes, n * 8, cudaMemcpyHostToDevice));
    GNB_CUDA(cudaMemcpy(d_b, bins, n * 4, cudaMemcpyHostToDevice));
    launch_emplace(t.d_data, t.bin_size, (uint32_t)t.hash_shift, (uint32_t)t.hash_funs, (uint32_t)t.row_words(), (uint32_t)t.w0, d_h, d_b, n, 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(d_h);
    cudaFree(d_b);
    GNB_CUDA(e);
    return GNB_OK;
}

extern "C" int gnb_db_set_targets(gnb_db *db, uint64_t n_targets, const char *const *names, const uint32_t *bin_target,
                                  const uint64_t *target_hashes, uint64_t max_hashes_bin)
{
    if (!db || db->is_hibf || !names || !bin_target || !target_hashes || n_targets == 0)
        return fail(GNB_ERR_ARG, "gnb_db_set_targets: bad arguments");
    const IbfHost &t = db->ibfs[0];
    for (uint64_t b = 0; b < t.bins; ++b)
        if (bin_target[b] >= n_targets)
            return fail(GNB_ERR_ARG, "gnb_db_set_targets: target index out of range");
    db->bin_map.clear();
    db->hashes_count.clear();
    for (uint64_t b = 0; b < t.bins; ++b)
        db->bin_map.emplace_back(b, names[bin_target[b]]);
    for (uint64_t i = 0; i < n_targets; ++i)
        db->hashes_count.emplace_back(names[i], target_hashes[i]);
    db->max_hashes_bin = max_hashes_bin;
    db->derive_targets();
    return GNB_OK;
}

extern "C" int gnb_db_set_fp(gnb_db *db, double max_fp, double true_max_fp, double true_avg_fp)
{
    if (!db || db->is_hibf)
        return fail(GNB_ERR_ARG, "gnb_db_set_fp: bad arguments");
    db->max_fp      = max_fp;
    db->true_max_fp = true_max_fp;
    db->true_avg_fp = true_avg_fp;
    db->derive_targets();
    return GNB_OK;
}

extern "C" int gnb_db_read_words(const gnb_db *db, uint64_t ibf_index, uint64_t word_offset, uint64_t n_words, uint64_t *out)
{
    if (!db || ibf_index >= db->ibfs.size() || !out)
        return fail(GNB_ERR_ARG, "gnb_db_read_words: bad arguments");
    if (db->ibfs[0].paged())
        return fail(GNB_ERR_ARG, "not available on a paged filter (host-resident tier): the bitvector is not whole in HBM");
    const IbfHost &t = db->ibfs[ibf_index];
    if (word_offset + n_words > t.bin_size * t.row_words())
        return fail(GNB_ERR_ARG, "gnb_db_read_words: range out of bounds");
    GNB_CUDA(cudaSetDevice(db->device));
    GNB_CUDA(cudaMemcpy(out, t.d_data + word_offset, n_words * 8, cudaMemcpyDeviceToHost));
    return GNB_OK;
}

namespace
{
// seqan3 IBF serialize (IBF.hpp:561-571) + sdsl::bit_vector (int_vector.hpp:2029-2035), bitvector streamed from HBM
bool write_ibf_body(FILE *fp, const gnb::IbfHost &t, std::vector<uint64_t> &buf)
{
    auto put = [&](const void *p, size_t n) { return fwrite(p, 1, n, fp) == n; };
    bool ok  = put(&t.bins, 8) && put(&t.technical_bins, 8) && put(&t.bin_size, 8) && put(&t.hash_shift, 8) && put(&t.bin_words, 8) && put(&t.hash_funs, 8);
    uint8_t  width  = 1;
    float    growth = 1.5f;
    uint64_t n_bits = t.technical_bins * t.bin_size;
    ok &= put(&width, 1) && put(&growth, 4) && put(&n_bits, 8);
    const uint64_t total = t.bin_size * t.bin_words;
    const uint64_t step  = 8u << 20; // words
    buf.resize(std::min(step, std::max<uint64_t>(total, 1)));
    for (uint64_t o = 0; ok && o < total; o += step)
    {
        const uint64_t m = std::min(step, total - o);
        if (cudaMemcpy(buf.data(), t.d_data + o, m * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
            return false;
        ok &= put(buf.data(), m * 8);
    }
    return ok;
}
} // namespace

// raptor 3.0.1 index layout as read by load_filter(THIBF) GC.cpp:875-938 (see gnb_db_open)
static int save_hibf(const gnb_db *db, const char *path)
{
    FILE *fp = fopen(path, "wb");
    if (!fp)
        return fail(GNB_ERR_IO, std::string("cannot write ") + path);
    auto put = [&](const void *p, size_t n) { return fwrite(p, 1, n, fp) == n; };
    auto put_str = [&](const std::string &x) {
        uint64_t l = x.size();
        return put(&l, 8) && put(x.data(), l);
    };
    uint32_t version = 1;
    uint64_t window = db->window_size, shape_size = db->kmer_size, shape_bits = db->kmer_size >= 64 ? ~0ull : ((1ull << db->kmer_size) - 1);
    uint8_t  parts = 1, compressed = 0, is_hibf = 1;
    bool     ok = put(&version, 4) && put(&window, 8) && put(&shape_size, 8) && put(&shape_bits, 8) && put(&parts, 1) && put(&compressed, 1);
    // bin_path: one file per user bin, named after the target
    std::vector<std::string> names(db->n_user_bins);
    for (auto const &[u, name] : db->bin_map)
        if (u < names.size() && names[u].empty())
            names[u] = name;
    uint64_t n = names.size();
    ok &= put(&n, 8);
    for (auto const &nm : names)
    {
        uint64_t one = 1;
        ok &= put(&one, 8) && put_str("/db/" + nm + ".minimiser");
    }
    ok &= put(&db->max_fp, 8) && put(&is_hibf, 1);
    n = db->ibfs.size();
    ok &= put(&n, 8);
    cudaSetDevice(db->device);
    std::vector<uint64_t> buf;
    for (size_t i = 0; ok && i < db->ibfs.size(); ++i)
        ok &= write_ibf_body(fp, db->ibfs[i], buf);
    auto put_vv = [&](const std::vector<std::vector<int64_t>> &vv) {
        uint64_t m = vv.size();
        bool     o = put(&m, 8);
        for (auto const &v : vv)
        {
            uint64_t l = v.size();
            o &= put(&l, 8) && (l == 0 || put(v.data(), l * 8));
        }
        return o;
    };
    ok &= put_vv(db->next_ibf_id);
    n = names.size();
    ok &= put(&n, 8);
    for (auto const &nm : names)
        ok &= put_str(nm);
    ok &= put_vv(db->bin_to_user);
    ok &= fclose(fp) == 0;
    return ok ? GNB_OK : fail(GNB_ERR_IO, std::string("short write to ") + path);
}

extern "C" int gnb_db_create_hibf(uint64_t n_ibfs, const uint64_t *bins, const uint64_t *bin_size_bits, uint32_t hash_functions, uint32_t kmer_size,
                                  uint32_t window_size, const int64_t *next_ibf, const int64_t *bin_to_user, uint64_t n_user_bins,
                                  const char *const *user_bin_names, double fpr, int device, gnb_db **out)
{
    if (!out || n_ibfs == 0 || !bins || !bin_size_bits || !next_ibf || !bin_to_user || !user_bin_names || n_user_bins == 0 || hash_functions < 1 ||
        hash_functions > 5 || kmer_size < 1 || kmer_size > 32 || window_size < kmer_size)
        return fail(GNB_ERR_ARG, "gnb_db_create_hibf: bad arguments");
    GNB_CUDA(cudaSetDevice(device));
    std::unique_ptr<gnb_db> db(new gnb_db);
    db->is_hibf     = true;
    db->device      = device;
    db->kmer_size   = kmer_size;
    db->window_size = window_size;
    db->max_fp      = fpr;
    db->n_user_bins = n_user_bins;
    db->ibfs.resize(n_ibfs);
    db->next_ibf_id.resize(n_ibfs);
    db->bin_to_user.resize(n_ibfs);
    uint64_t off = 0;
    int      rc  = GNB_OK;
    for (uint64_t i = 0; i < n_ibfs && rc == GNB_OK; ++i)
    {
        IbfHost &t = db->ibfs[i];
        if (bins[i] == 0 || bin_size_bits[i] == 0)
        {
            rc = fail(GNB_ERR_ARG, "gnb_db_create_hibf: empty sub-IBF");
            break;
        }
        t.bins           = bins[i];
        t.bin_words      = (bins[i] + 63) >> 6;
        t.technical_bins = t.bin_words << 6;
        t.bin_size       = bin_size_bits[i];
        t.hash_shift     = (uint64_t)__builtin_clzll(bin_size_bits[i]);
        t.hash_funs      = hash_functions;
        t.w0             = 0;
        t.w1             = t.bin_words;
        db->next_ibf_id[i].assign(next_ibf + off, next_ibf + off + bins[i]);
        db->bin_to_user[i].assign(bin_to_user + off, bin_to_user + off + bins[i]);
        for (uint64_t b = 0; b < bins[i]; ++b)
        {
            const int64_t fi = bin_to_user[off + b], nx = next_ibf[off + b];
            if (fi >= (int64_t)n_user_bins || (fi < 0 && (nx < 0 || nx >= (int64_t)n_ibfs)))
                rc = fail(GNB_ERR_ARG, "gnb_db_create_hibf: table entry out of range");
        }
        off += bins[i];
        if (rc == GNB_OK && cudaMalloc((void **)&t.d_data, t.device_bytes()) != cudaSuccess)
            rc = fail(GNB_ERR_CUDA, "gnb_db_create_hibf: out of device memory");
        if (rc == GNB_OK)
            cudaMemsetAsync(t.d_data, 0, t.device_bytes(), 0);
    }
    if (rc == GNB_OK && cudaDeviceSynchronize() != cudaSuccess)
        rc = fail(GNB_ERR_CUDA, "gnb_db_create_hibf: memset failed");
    if (rc != GNB_OK)
    {
        std::string keep = g_err;
        gnb_db_free(db.release());
        g_err = keep;
        return rc;
    }
    for (uint64_t u = 0; u < n_user_bins; ++u)
        db->bin_map.emplace_back(u, user_bin_names[u] ? user_bin_names[u] : "");
    db->derive_targets();
    *out = db.release();
    return GNB_OK;
}

extern "C" int gnb_db_save(const gnb_db *db, const char *path)
{
    if (!db || !path)
        return fail(GNB_ERR_ARG, "gnb_db_save: bad arguments");
    if (db->is_hibf)
        return save_hibf(db, path);
    const IbfHost &t = db->ibfs[0];
    if (t.w0 != 0 || t.w1 != t.bin_words)
        return fail(GNB_ERR_ARG, "gnb_db_save: sharded handle");
    if (db->ibfs[0].paged())
        return fail(GNB_ERR_ARG, "not available on a paged filter (host-resident tier): the bitvector is not whole in HBM");
    FILE *fp = fopen(path, "wb");
    if (!fp)
        return fail(GNB_ERR_IO, std::string("cannot write ") + path);
    auto put = [&](const void *p, size_t n) { return fwrite(p, 1, n, fp) == n; };
    bool ok  = true;
    for (int i = 0; i < 3; ++i)
    {
        int32_t v = db->version[i];
        ok &= put(&v, 4);
    }
    uint8_t  hf = (uint8_t)t.hash_funs, k = (uint8_t)db->kmer_size;
    uint16_t w  = (uint16_t)db->window_size;
    ok &= put(&t.bins, 8) && put(&db->max_hashes_bin, 8) && put(&hf, 1) && put(&k, 1) && put(&w, 2) && put(&t.bin_size, 8) &&
          put(&db->max_fp, 8) && put(&db->true_max_fp, 8) && put(&db->true_avg_fp, 8);
    uint64_t n = db->hashes_count.size();
    ok &= put(&n, 8);
    for (auto const &[name, c] : db->hashes_count)
    {
        uint64_t l = name.size();
        ok &= put(&l, 8) && put(name.data(), l) && put(&c, 8);
    }
    n = db->bin_map.size();
    ok &= put(&n, 8);
    for (auto const &[b, name] : db->bin_map)
    {
        uint64_t l = name.size();
        ok &= put(&b, 8) && put(&l, 8) && put(name.data(), l);
    }
    cudaSetDevice(db->device);
    std::vector<uint64_t> buf;
    ok &= write_ibf_body(fp, t, buf);
    ok &= fclose(fp) == 0;
    return ok ? GNB_OK : fail(GNB_ERR_IO, std::string("short write to ") + path);
}
