// Host-side knowledge-graph index builder + the error/buffer plumbing of libmre_b200.so.
//
// What it replaces in the reference (paths relative to /root/reference):
//   OpenKE/openke/base/Reader.h:53-160   importTrainFiles: de-duplicated train list in (h,r,t) and (t,r,h)
//                                        order, tph/hpt per relation
//   OpenKE/openke/base/Reader.h:167-257  importTestFiles: all-splits membership list, test/valid in (r,h,t) order
//   OpenKE/openke/base/Corrupt.h:166-177 _find
// The reference keeps per-entity [lef,rig] range tables and array-of-struct triples for pointer-chasing
// binary searches on the CPU.  Here every table is a struct-of-arrays with ONE packed int64 sort key per row
// (entity * R + relation), so a (entity, relation) run is found on the GPU with two lower_bounds on a flat,
// coalescable key column, and the run's payload column (the known tails / heads) is contiguous and sorted.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <numeric>
#include <string>

#include "common.h"

namespace mre {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return MRE_OK;
    if (p) {
        MRE_CUDA(cudaDeviceSynchronize());
        MRE_CUDA(cudaFree(p));
        p = nullptr;
        cap = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    MRE_CUDA(cudaMalloc(&p, want));
    cap = want;
    return MRE_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}
int PinnedBuf::reserve(size_t bytes) {
    if (bytes <= cap) return MRE_OK;
    if (p) {
        MRE_CUDA(cudaDeviceSynchronize());
        MRE_CUDA(cudaFreeHost(p));
        p = nullptr;
        cap = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    MRE_CUDA(cudaMallocHost(&p, want));
    cap = want;
    return MRE_OK;
}
void PinnedBuf::release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
}

static inline bool less_hrt(const Triple &a, const Triple &b) {
    if (a.h != b.h) return a.h < b.h;
    if (a.r != b.r) return a.r < b.r;
    return a.t < b.t;
}
static inline bool less_trh(const Triple &a, const Triple &b) {
    if (a.t != b.t) return a.t < b.t;
    if (a.r != b.r) return a.r < b.r;
    return a.h < b.h;
}
static inline bool less_rht(const Triple &a, const Triple &b) {
    if (a.r != b.r) return a.r < b.r;
    if (a.h != b.h) return a.h < b.h;
    return a.t < b.t;
}
static inline bool same(const Triple &a, const Triple &b) { return a.h == b.h && a.r == b.r && a.t == b.t; }

static int pack(const int64_t *h, const int64_t *t, const int64_t *r, int64_t n, int64_t E, int64_t R, const char *what,
                std::vector<Triple> &out) {
    out.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        if (h[i] < 0 || h[i] >= E || t[i] < 0 || t[i] >= E || r[i] < 0 || r[i] >= R) {
            set_error("%s triple %lld = (h=%lld, t=%lld, r=%lld) out of range for E=%lld, R=%lld", what, (long long)i,
                      (long long)h[i], (long long)t[i], (long long)r[i], (long long)E, (long long)R);
            return MRE_ERR_INVALID;
        }
        out[(size_t)i] = Triple{h[i], r[i], t[i]};
    }
    return MRE_OK;
}

// std::sort over `threads` chunks on as many host threads, then pairwise std::inplace_merge rounds (the chunks of a round merge
// in parallel too).  The reference sorts its five lists one after the other on one thread (Reader.h:107-109,201-227).
template <class Less>
static void parallel_sort(std::vector<Triple> &v, Less less, int threads) {
    const size_t n = v.size();
    if (threads <= 1 || n < 65536) {
        std::sort(v.begin(), v.end(), less);
        return;
    }
    std::vector<size_t> cut((size_t)threads + 1);
    for (int i = 0; i <= threads; i++) cut[(size_t)i] = n * (size_t)i / (size_t)threads;
    {
        std::vector<std::thread> pool;
        for (int i = 0; i < threads; i++)
            pool.emplace_back([&, i] { std::sort(v.begin() + (ptrdiff_t)cut[(size_t)i], v.begin() + (ptrdiff_t)cut[(size_t)i + 1], less); });
        for (auto &t : pool) t.join();
    }
    for (int width = 1; width < threads; width *= 2) {
        std::vector<std::thread> pool;
        for (int i = 0; i + width < threads; i += 2 * width) {
            const size_t lo = cut[(size_t)i], mid = cut[(size_t)(i + width)], hi = cut[(size_t)std::min(i + 2 * width, threads)];
            pool.emplace_back([&, lo, mid, hi] { std::inplace_merge(v.begin() + (ptrdiff_t)lo, v.begin() + (ptrdiff_t)mid, v.begin() + (ptrdiff_t)hi, less); });
        }
        for (auto &t : pool) t.join();
    }
}

static int host_threads() {
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(hc ? hc : 1u, 16u));
}

static int build(mre_index *ix, std::vector<Triple> &train_raw, std::vector<Triple> &valid, std::vector<Triple> &test) {
    const int64_t R = ix->R;
    ix->n_train_raw = (int64_t)train_raw.size();
    // the three independent chains -- membership list, the two train orders, the test / valid orders -- run side by side, each
    // sort itself split over a share of the host threads
    const int per = std::max(1, host_threads() / 3);

    // membership list: test + RAW train + valid, (h,r,t) order, duplicates kept (Reader.h:201-226)
    ix->all_head.reserve(test.size() + train_raw.size() + valid.size());
    ix->all_head.insert(ix->all_head.end(), test.begin(), test.end());
    ix->all_head.insert(ix->all_head.end(), train_raw.begin(), train_raw.end());
    ix->all_head.insert(ix->all_head.end(), valid.begin(), valid.end());
    std::thread chain_all([&] { parallel_sort(ix->all_head, less_hrt, per); });
    std::thread chain_test([&] {
        parallel_sort(test, less_rht, std::max(1, per / 2));
        parallel_sort(valid, less_rht, std::max(1, per / 2));
    });

    // train: sort, drop duplicates (Reader.h:91-105), second order (Reader.h:107-109)
    parallel_sort(train_raw, less_hrt, per);
    train_raw.erase(std::unique(train_raw.begin(), train_raw.end(), same), train_raw.end());
    ix->train_head.swap(train_raw);
    ix->train_tail = ix->train_head;
    parallel_sort(ix->train_tail, less_trh, per);
    chain_all.join();
    chain_test.join();

    // tph / hpt in float32 exactly as Reader.h:142-159: float counters of distinct (h,r) / (t,r) pairs,
    // then (integer frequency) / (float count)
    std::vector<int64_t> freq((size_t)R, 0);
    ix->left_mean.assign((size_t)R, 0.f);
    ix->right_mean.assign((size_t)R, 0.f);
    const size_t n = ix->train_head.size();
    for (size_t i = 0; i < n; i++) {
        const Triple &a = ix->train_head[i];
        freq[(size_t)a.r]++;
        if (i == 0 || a.h != ix->train_head[i - 1].h || a.r != ix->train_head[i - 1].r) ix->left_mean[(size_t)a.r] += 1.0f;
        const Triple &b = ix->train_tail[i];
        if (i == 0 || b.t != ix->train_tail[i - 1].t || b.r != ix->train_tail[i - 1].r) ix->right_mean[(size_t)b.r] += 1.0f;
    }
    ix->bern_prob.assign((size_t)R, 500.f);
    for (int64_t r = 0; r < R; r++) {
        ix->left_mean[(size_t)r] = (float)freq[(size_t)r] / ix->left_mean[(size_t)r];
        ix->right_mean[(size_t)r] = (float)freq[(size_t)r] / ix->right_mean[(size_t)r];
        // Base.cpp:113: prob = 1000 * right_mean / (right_mean + left_mean), all REAL
        volatile float num = 1000 * ix->right_mean[(size_t)r];
        volatile float den = ix->right_mean[(size_t)r] + ix->left_mean[(size_t)r];
        ix->bern_prob[(size_t)r] = num / den;
    }

    ix->test.swap(test);
    ix->valid.swap(valid);
    return MRE_OK;
}

static int read_count(const std::string &path, int64_t *out) {
    FILE *f = fopen(path.c_str(), "r");
    if (!f) {
        set_error("cannot open %s", path.c_str());
        return MRE_ERR_IO;
    }
    long v = 0;
    int ok = fscanf(f, "%ld", &v);
    fclose(f);
    if (ok != 1) {
        set_error("%s: missing count line", path.c_str());
        return MRE_ERR_IO;
    }
    *out = v;
    return MRE_OK;
}

// "<count>\n" then count rows "h t r" (OpenKE/README.md:126-141; Reader.h:83-88 reads h, t, r in that order)
static int read_triples(const std::string &path, bool optional, int64_t E, int64_t R, std::vector<Triple> &out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
        if (optional) return MRE_OK;
        set_error("cannot open %s", path.c_str());
        return MRE_ERR_IO;
    }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::string buf((size_t)sz + 1, '\0');
    size_t got = fread(&buf[0], 1, (size_t)sz, f);
    fclose(f);
    buf[got] = '\0';
    const char *p = buf.c_str();
    char *end = nullptr;
    long long n = strtoll(p, &end, 10);
    if (end == p || n < 0) {
        set_error("%s: missing count line", path.c_str());
        return MRE_ERR_IO;
    }
    p = end;
    out.resize((size_t)n);
    for (long long i = 0; i < n; i++) {
        long long v[3];
        for (int k = 0; k < 3; k++) {
            v[k] = strtoll(p, &end, 10);
            if (end == p) {
                set_error("%s: truncated at triple %lld of %lld", path.c_str(), i, n);
                return MRE_ERR_IO;
            }
            p = end;
        }
        if (v[0] < 0 || v[0] >= E || v[1] < 0 || v[1] >= E || v[2] < 0 || v[2] >= R) {
            set_error("%s: triple %lld = (h=%lld, t=%lld, r=%lld) out of range for E=%lld, R=%lld", path.c_str(), i, v[0],
                      v[1], v[2], (long long)E, (long long)R);
            return MRE_ERR_INVALID;
        }
        out[(size_t)i] = Triple{v[0], v[2], v[1]};
    }
    return MRE_OK;
}

// type_constrain.txt (Reader.h:267-317; written by benchmarks/*/n-n.py): a count line, then per relation two rows
// "rel n id*n" -- the admissible heads, then the admissible tails.  Lists are kept sorted and de-duplicated: the reference
// sorts them (:299,:307) and its merge pointer (Test.h:89-90) lets a repeated id count once.
static int read_type_constrain(const std::string &path, mre_index *ix) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
        set_error("cannot open %s", path.c_str());
        return MRE_ERR_IO;
    }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::string buf((size_t)sz + 1, '\0');
    size_t got = fread(&buf[0], 1, (size_t)sz, f);
    fclose(f);
    buf[got] = '\0';
    const char *p = buf.c_str();
    char *end = nullptr;
    auto next = [&](long long *v) {
        *v = strtoll(p, &end, 10);
        if (end == p) return false;
        p = end;
        return true;
    };
    long long n_rel = 0;
    if (!next(&n_rel) || n_rel < 0) {
        set_error("%s: missing count line", path.c_str());
        return MRE_ERR_IO;
    }
    std::vector<std::vector<int64_t>> lists[2];
    lists[0].resize((size_t)ix->R);
    lists[1].resize((size_t)ix->R);
    for (long long i = 0; i < n_rel; i++) {
        for (int side = 0; side < 2; side++) {
            long long rel = 0, tot = 0;
            if (!next(&rel) || !next(&tot) || tot < 0) {
                set_error("%s: truncated at relation row %lld", path.c_str(), 2 * i + side);
                return MRE_ERR_IO;
            }
            if (rel < 0 || rel >= ix->R) {
                set_error("%s: relation %lld out of range for R=%lld", path.c_str(), rel, (long long)ix->R);
                return MRE_ERR_INVALID;
            }
            std::vector<int64_t> &dst = lists[side][(size_t)rel];
            dst.clear();
            dst.reserve((size_t)tot);
            for (long long j = 0; j < tot; j++) {
                long long e = 0;
                if (!next(&e)) {
                    set_error("%s: relation %lld: list shorter than its count %lld", path.c_str(), rel, tot);
                    return MRE_ERR_IO;
                }
                if (e < 0 || e >= ix->E) {
                    set_error("%s: relation %lld: entity %lld out of range for E=%lld", path.c_str(), rel, e, (long long)ix->E);
                    return MRE_ERR_INVALID;
                }
                dst.push_back(e);
            }
        }
    }
    for (int side = 0; side < 2; side++) {
        ix->type_ptr[side].assign((size_t)ix->R + 1, 0);
        ix->type_idx[side].clear();
        for (int64_t r = 0; r < ix->R; r++) {
            std::vector<int64_t> &l = lists[side][(size_t)r];
            std::sort(l.begin(), l.end());
            l.erase(std::unique(l.begin(), l.end()), l.end());
            ix->type_idx[side].insert(ix->type_idx[side].end(), l.begin(), l.end());
            ix->type_ptr[side][(size_t)r + 1] = (int64_t)ix->type_idx[side].size();
        }
    }
    ix->has_type = true;
    return MRE_OK;
}

// entity2id.txt / relation2id.txt counts + the three *2id.txt splits of a benchmark directory (`dir` ends with '/')
int read_benchmark_dir(const std::string &dir, int64_t *E, int64_t *R, std::vector<Triple> &tr, std::vector<Triple> &va, std::vector<Triple> &te) {
    MRE_TRY(read_count(dir + "entity2id.txt", E));
    MRE_TRY(read_count(dir + "relation2id.txt", R));
    MRE_CHECK_ARG(*E > 0 && *R > 0, "%s: E and R must be positive", dir.c_str());
    MRE_TRY(read_triples(dir + "train2id.txt", false, *E, *R, tr));
    MRE_TRY(read_triples(dir + "valid2id.txt", true, *E, *R, va));
    MRE_TRY(read_triples(dir + "test2id.txt", true, *E, *R, te));
    return MRE_OK;
}

template <class T>
static int upload(const std::vector<T> &v, T **dst) {
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    MRE_CUDA(cudaMalloc((void **)dst, bytes));
    if (!v.empty()) MRE_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return MRE_OK;
}

}  // namespace mre

using namespace mre;

extern "C" {

const char *mre_last_error(void) { return g_err; }
int mre_abi_version(void) { return MRE_ABI_VERSION; }

int mre_device_ok(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= device || device < 0) {
        set_error("no CUDA device %d (%s)", device, e == cudaSuccess ? "count too small" : cudaGetErrorString(e));
        return MRE_ERR_CUDA;
    }
    cudaDeviceProp p;
    MRE_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        set_error("device %d is sm_%d%d; libmre_b200 is built for sm_100a only", device, p.major, p.minor);
        return MRE_ERR_CUDA;
    }
    return MRE_OK;
}

int mre_index_create(int64_t E, int64_t R, const int64_t *train_h, const int64_t *train_t, const int64_t *train_r,
                     int64_t n_train, const int64_t *valid_h, const int64_t *valid_t, const int64_t *valid_r,
                     int64_t n_valid, const int64_t *test_h, const int64_t *test_t, const int64_t *test_r, int64_t n_test,
                     mre_index **out) {
    MRE_CHECK_ARG(out != nullptr, "out is NULL");
    MRE_CHECK_ARG(E > 0 && R > 0, "E and R must be positive");
    MRE_CHECK_ARG(n_train >= 0 && n_valid >= 0 && n_test >= 0, "negative split size");
    MRE_CHECK_ARG(E <= (INT64_MAX / 2) / R, "E * R overflows the packed key");
    mre_index *ix = new mre_index();
    ix->E = E;
    ix->R = R;
    std::vector<Triple> tr, va, te;
    int rc = pack(train_h, train_t, train_r, n_train, E, R, "train", tr);
    if (rc == MRE_OK) rc = pack(valid_h, valid_t, valid_r, n_valid, E, R, "valid", va);
    if (rc == MRE_OK) rc = pack(test_h, test_t, test_r, n_test, E, R, "test", te);
    if (rc == MRE_OK) rc = build(ix, tr, va, te);
    if (rc != MRE_OK) {
        delete ix;
        return rc;
    }
    *out = ix;
    return MRE_OK;
}

int mre_index_create_from_dir(const char *in_path, mre_index **out) {
    MRE_CHECK_ARG(in_path && out, "NULL argument");
    std::string dir(in_path);
    if (!dir.empty() && dir.back() != '/') dir += '/';
    int64_t E = 0, R = 0;
    std::vector<Triple> tr, va, te;
    MRE_TRY(read_benchmark_dir(dir, &E, &R, tr, va, te));
    mre_index *ix = new mre_index();
    ix->E = E;
    ix->R = R;
    int rc = build(ix, tr, va, te);
    if (rc == MRE_OK) {   // importTypeFiles (Reader.h:267-317); the file is optional here, the reference crashes without it
        FILE *f = fopen((dir + "type_constrain.txt").c_str(), "r");
        if (f) {
            fclose(f);
            rc = read_type_constrain(dir + "type_constrain.txt", ix);
        }
    }
    if (rc != MRE_OK) {
        delete ix;
        return rc;
    }
    *out = ix;
    return MRE_OK;
}

void mre_index_destroy(mre_index *ix) {
    if (!ix) return;
    if (ix->device >= 0) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(ix->device);
        cudaFree(ix->d_all_hr_key); cudaFree(ix->d_all_hr_val); cudaFree(ix->d_all_tr_key); cudaFree(ix->d_all_tr_val);
        cudaFree(ix->d_tr_h); cudaFree(ix->d_tr_r); cudaFree(ix->d_tr_t); cudaFree(ix->d_tr_hr_key);
        cudaFree(ix->d_tr_tr_key); cudaFree(ix->d_tr_tr_val); cudaFree(ix->d_bern_prob);
        for (int side = 0; side < 2; side++) { cudaFree(ix->d_type_ptr[side]); cudaFree(ix->d_type_idx[side]); }
        cudaSetDevice(cur);
    }
    delete ix;
}

// the type-constraint lists follow the index onto its device (corrupt(), Corrupt.h:179-195, draws from them there)
static int upload_type_lists(mre_index *ix) {
    if (ix->device < 0 || !ix->has_type) return MRE_OK;
    MRE_CUDA(cudaSetDevice(ix->device));
    for (int side = 0; side < 2; side++) {
        if (ix->d_type_ptr[side]) { cudaFree(ix->d_type_ptr[side]); ix->d_type_ptr[side] = nullptr; }
        if (ix->d_type_idx[side]) { cudaFree(ix->d_type_idx[side]); ix->d_type_idx[side] = nullptr; }
        std::vector<int64_t> idx = ix->type_idx[side];
        if (idx.empty()) idx.push_back(0);               // keep the pointer non-NULL
        MRE_TRY(upload(ix->type_ptr[side], &ix->d_type_ptr[side]));
        MRE_TRY(upload(idx, &ix->d_type_idx[side]));
    }
    return MRE_OK;
}

int mre_index_to_device(mre_index *ix, int device) {
    MRE_CHECK_ARG(ix != nullptr, "index is NULL");
    if (ix->device == device) return MRE_OK;
    MRE_CHECK_ARG(ix->device < 0, "index already lives on device %d", ix->device);
    MRE_TRY(mre_device_ok(device));
    MRE_CUDA(cudaSetDevice(device));
    const int64_t R = ix->R;
    // filter tables: all splits, de-duplicated, as (key, payload) columns in both orientations
    std::vector<Triple> all = ix->all_head;
    all.erase(std::unique(all.begin(), all.end(), same), all.end());
    std::vector<int64_t> key(all.size()), val(all.size());
    for (size_t i = 0; i < all.size(); i++) { key[i] = all[i].h * R + all[i].r; val[i] = all[i].t; }
    MRE_TRY(upload(key, &ix->d_all_hr_key));
    MRE_TRY(upload(val, &ix->d_all_hr_val));
    std::sort(all.begin(), all.end(), less_trh);
    for (size_t i = 0; i < all.size(); i++) { key[i] = all[i].t * R + all[i].r; val[i] = all[i].h; }
    MRE_TRY(upload(key, &ix->d_all_tr_key));
    MRE_TRY(upload(val, &ix->d_all_tr_val));
    ix->n_all = (int64_t)all.size();
    // sampler tables: de-duplicated train
    const size_t n = ix->train_head.size();
    std::vector<int64_t> ch(n), cr(n), ct(n), k1(n), k2(n), v2(n);
    for (size_t i = 0; i < n; i++) {
        const Triple &a = ix->train_head[i];
        ch[i] = a.h; cr[i] = a.r; ct[i] = a.t; k1[i] = a.h * R + a.r;
        const Triple &b = ix->train_tail[i];
        k2[i] = b.t * R + b.r; v2[i] = b.h;
    }
    MRE_TRY(upload(ch, &ix->d_tr_h));
    MRE_TRY(upload(cr, &ix->d_tr_r));
    MRE_TRY(upload(ct, &ix->d_tr_t));
    MRE_TRY(upload(k1, &ix->d_tr_hr_key));
    MRE_TRY(upload(k2, &ix->d_tr_tr_key));
    MRE_TRY(upload(v2, &ix->d_tr_tr_val));
    MRE_TRY(upload(ix->bern_prob, &ix->d_bern_prob));
    ix->n_train = (int64_t)n;
    ix->device = device;
    return upload_type_lists(ix);
}

int64_t mre_index_device_column(const mre_index *ix, int which, void *host_out) {
    if (!ix || ix->device < 0) {
        set_error("the index has no device tables (call mre_index_to_device or build with mre_index_create_device)");
        return -1;
    }
    const void *src[11] = {ix->d_all_hr_key, ix->d_all_hr_val, ix->d_all_tr_key, ix->d_all_tr_val, ix->d_tr_h, ix->d_tr_r, ix->d_tr_t,
                           ix->d_tr_hr_key, ix->d_tr_tr_key, ix->d_tr_tr_val, ix->d_bern_prob};
    if (which < 0 || which > 10) {
        set_error("unknown device column %d", which);
        return -1;
    }
    const int64_t n = which < 4 ? ix->n_all : which < 10 ? ix->n_train : ix->R;
    if (host_out && n > 0) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(ix->device);
        const cudaError_t e = cudaMemcpy(host_out, src[which], (size_t)n * (which == 10 ? sizeof(float) : sizeof(int64_t)), cudaMemcpyDeviceToHost);
        cudaSetDevice(cur);
        if (e != cudaSuccess) {
            set_error("mre_index_device_column: %s", cudaGetErrorString(e));
            return -1;
        }
    }
    return n;
}

int64_t mre_index_total(const mre_index *ix, int which) {
    if (!ix) return -1;
    switch (which) {
        case MRE_TOTAL_ENTITY: return ix->E;
        case MRE_TOTAL_RELATION: return ix->R;
        case MRE_TOTAL_TRAIN: return (int64_t)ix->train_head.size();
        case MRE_TOTAL_VALID: return (int64_t)ix->valid.size();
        case MRE_TOTAL_TEST: return (int64_t)ix->test.size();
        case MRE_TOTAL_TRIPLE: return (int64_t)ix->all_head.size();
    }
    return -1;
}

int mre_index_get_split(const mre_index *ix, int split, int64_t *h, int64_t *t, int64_t *r) {
    MRE_CHECK_ARG(ix && h && t && r, "NULL argument");
    const std::vector<Triple> *v = split == MRE_SPLIT_TRAIN ? &ix->train_head
                                 : split == MRE_SPLIT_VALID ? &ix->valid
                                 : split == MRE_SPLIT_TEST  ? &ix->test : nullptr;
    MRE_CHECK_ARG(v != nullptr, "unknown split %d", split);
    for (size_t i = 0; i < v->size(); i++) { h[i] = (*v)[i].h; t[i] = (*v)[i].t; r[i] = (*v)[i].r; }
    return MRE_OK;
}

int mre_index_get_means(const mre_index *ix, float *tph, float *hpt) {
    MRE_CHECK_ARG(ix && tph && hpt, "NULL argument");
    memcpy(tph, ix->left_mean.data(), ix->left_mean.size() * sizeof(float));
    memcpy(hpt, ix->right_mean.data(), ix->right_mean.size() * sizeof(float));
    return MRE_OK;
}

static int upload_type_lists(mre_index *ix);

int mre_index_load_type_constrain(mre_index *ix, const char *path) {
    MRE_CHECK_ARG(ix && path, "NULL argument");
    MRE_TRY(read_type_constrain(path, ix));
    return upload_type_lists(ix);                       // no-op while the index is host-only
}

int mre_index_set_type_constrain(mre_index *ix, const int64_t *head_ptr, const int64_t *head_idx, const int64_t *tail_ptr,
                                 const int64_t *tail_idx) {
    MRE_CHECK_ARG(ix && head_ptr && tail_ptr, "NULL argument");
    const int64_t *ptrs[2] = {head_ptr, tail_ptr}, *idxs[2] = {head_idx, tail_idx};
    std::vector<int64_t> new_ptr[2], new_idx[2];
    for (int side = 0; side < 2; side++) {
        MRE_CHECK_ARG(ptrs[side][0] == 0, "type-constraint prefix must start at 0");
        new_ptr[side].assign((size_t)ix->R + 1, 0);
        for (int64_t r = 0; r < ix->R; r++) {
            const int64_t lo = ptrs[side][r], hi = ptrs[side][r + 1];
            MRE_CHECK_ARG(hi >= lo, "type-constraint prefix decreases at relation %lld", (long long)r);
            MRE_CHECK_ARG(hi == lo || idxs[side], "type-constraint ids are NULL");
            std::vector<int64_t> l(idxs[side] + lo, idxs[side] + hi);
            for (int64_t e : l) MRE_CHECK_ARG(e >= 0 && e < ix->E, "type-constraint entity %lld out of range", (long long)e);
            std::sort(l.begin(), l.end());
            l.erase(std::unique(l.begin(), l.end()), l.end());
            new_idx[side].insert(new_idx[side].end(), l.begin(), l.end());
            new_ptr[side][(size_t)r + 1] = (int64_t)new_idx[side].size();
        }
    }
    for (int side = 0; side < 2; side++) {
        ix->type_ptr[side].swap(new_ptr[side]);
        ix->type_idx[side].swap(new_idx[side]);
    }
    ix->has_type = true;
    return upload_type_lists(ix);
}

int64_t mre_index_type_total(const mre_index *ix, int side) {
    if (!ix || !ix->has_type || side < 0 || side > 1) return -1;
    return (int64_t)ix->type_idx[side].size();
}

int mre_index_get_type_constrain(const mre_index *ix, int side, int64_t *ptr, int64_t *idx) {
    MRE_CHECK_ARG(ix && ptr, "NULL argument");
    MRE_CHECK_ARG(side == 0 || side == 1, "side must be 0 or 1");
    MRE_CHECK_ARG(ix->has_type, "the index holds no type constraints (type_constrain.txt was not loaded)");
    memcpy(ptr, ix->type_ptr[side].data(), ix->type_ptr[side].size() * sizeof(int64_t));
    if (idx && !ix->type_idx[side].empty()) memcpy(idx, ix->type_idx[side].data(), ix->type_idx[side].size() * sizeof(int64_t));
    return MRE_OK;
}

int mre_index_find(const mre_index *ix, int64_t h, int64_t t, int64_t r) {
    if (!ix) return 0;
    Triple key{h, r, t};
    return std::binary_search(ix->all_head.begin(), ix->all_head.end(), key, less_hrt) ? 1 : 0;
}

int mre_ctx_create(int device, mre_ctx **out) {
    MRE_CHECK_ARG(out != nullptr, "out is NULL");
    MRE_TRY(mre_device_ok(device));
    MRE_CUDA(cudaSetDevice(device));
    mre_ctx *c = new mre_ctx();
    c->device = device;
    static std::atomic<int> next_slot{0};
    c->zsl_const_slot = next_slot.fetch_add(1) % 16;      // ZT_CONST_SLOTS (zsl_rank.cu)
    cudaDeviceProp p;
    MRE_CUDA(cudaGetDeviceProperties(&p, device));
    c->sm_count = p.multiProcessorCount;
    MRE_CUDA(cudaEventCreate(&c->ev0));
    MRE_CUDA(cudaEventCreate(&c->ev1));
    *out = c;
    return MRE_OK;
}

void mre_ctx_destroy(mre_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    mre::DevBuf *bufs[] = {&c->ent_n, &c->rel_n, &c->ent_aux, &c->ent_aux2, &c->qvec, &c->qvec2, &c->thr,
                           &c->tiles, &c->counters, &c->misc, &c->misc2, &c->stage_dev, &c->loss_acc, &c->stats, &c->met_scratch};
    for (auto *b : bufs) b->release();
    c->stage_pin.release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->aux) cudaStreamDestroy(c->aux);
    delete c;
}

int mre_ctx_sm_count(const mre_ctx *c) { return c ? c->sm_count : 0; }
int64_t mre_ctx_launch_count(const mre_ctx *c) { return c ? c->launches.load() : 0; }

}  // extern "C"
