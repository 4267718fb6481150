// multi.cuh -- several GPUs behind ONE caller: the multi-device context of include/phyfoo.h ("phyfoo_multi_*").
//
// A Java caller is one JVM process; SURVEY.md section 8b/8e ask for a library that takes a list of devices, shards the data
// set itself and owns the NCCL communicator.  Alignments are independent (section 8e): every device gets a contiguous block
// of alignments balanced by column count (the partition of phyfoo_b200/sharding.py::shard_bounds), the models are
// replicated, and the only exchange per consumer is a sum all-reduce of a few doubles -- [ll, w0, w1] per E-step,
// [f, f_1..f_dim] per gradient, the expected counts of the sufficient statistics -- enqueued with ncclAllReduce on each
// device's stream right behind the kernel that produces them, before the one device-to-host copy.
//
// One worker thread per device runs the per-device part of every call (the single-GPU entry points of this library are
// synchronous per context; a context belongs to one thread at a time, which is its worker here), so the devices work
// concurrently and each thread issues the collective for its own rank, which is NCCL's one-thread-per-device mode.
// NCCL is bound at run time (dlopen of libnccl.so.2; PHYFOO_NCCL_LIB overrides the name): the library has no link-time
// dependency on it and single-GPU users never load it.
#pragma once
#include <dlfcn.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

struct NcclApi {
    void* handle = nullptr;
    int (*CommInitAll)(void** comms, int ndev, const int* devlist) = nullptr;
    int (*CommDestroy)(void* comm) = nullptr;
    int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t stream) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    static const int kFloat64 = 8, kSum = 0;       // ncclFloat64, ncclSum (nccl.h)
    std::string load() {
        const char* names[3] = {getenv("PHYFOO_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) return "libnccl.so.2 not found (set PHYFOO_NCCL_LIB or LD_LIBRARY_PATH): " + std::string(dlerror() ? dlerror() : "");
        CommInitAll = (decltype(CommInitAll))dlsym(handle, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        if (!CommInitAll || !CommDestroy || !AllReduce || !GetErrorString) return "libnccl: a required symbol is missing";
        return "";
    }
};

// one thread per device; run(fn) executes fn(rank) on every worker and returns the first non-zero result
struct DeviceWorkers {
    struct Slot {
        std::thread th; std::mutex m; std::condition_variable cv;
        std::function<int(int)> job; bool has = false, done = false, quit = false; int rc = 0;
    };
    std::vector<Slot*> slots;
    void start(int n) {
        for (int r = 0; r < n; ++r) {
            Slot* s = new Slot();
            slots.push_back(s);
            s->th = std::thread([s, r]() {
                for (;;) {
                    std::unique_lock<std::mutex> lk(s->m);
                    s->cv.wait(lk, [s]() { return s->has || s->quit; });
                    if (s->quit) return;
                    std::function<int(int)> job = s->job;
                    lk.unlock();
                    const int rc = job(r);
                    lk.lock();
                    s->rc = rc; s->has = false; s->done = true;
                    s->cv.notify_all();
                }
            });
        }
    }
    int run(const std::function<int(int)>& fn) {
        for (Slot* s : slots) {
            std::lock_guard<std::mutex> lk(s->m);
            s->job = fn; s->has = true; s->done = false;
            s->cv.notify_all();
        }
        int first = 0;
        for (Slot* s : slots) {
            std::unique_lock<std::mutex> lk(s->m);
            s->cv.wait(lk, [s]() { return s->done; });
            if (s->rc && !first) first = s->rc;
        }
        return first;
    }
    void stop() {
        for (Slot* s : slots) {
            { std::lock_guard<std::mutex> lk(s->m); s->quit = true; s->cv.notify_all(); }
            s->th.join();
            delete s;
        }
        slots.clear();
    }
};

struct phyfoo_mctx {
    std::vector<phyfoo_ctx*> ctx;
    std::vector<void*> comm;
    std::vector<double*> d_red;          // per device: scratch for the all-reduced values
    size_t red_cap = 0;
    NcclApi nccl;
    DeviceWorkers workers;
    std::string err;
};
struct phyfoo_mdataset { std::vector<phyfoo_dataset*> ds; std::vector<int32_t> bounds; std::vector<int64_t> col_off; int S = 0; };
struct phyfoo_mmodel { std::vector<phyfoo_model*> m; };
struct phyfoo_mfeset { std::vector<phyfoo_feset*> fe; };

static thread_local std::string g_minit_err;

static int mfail(phyfoo_mctx* mc, int code, const std::string& msg) {
    if (mc) mc->err = msg; else g_minit_err = msg;
    return code;
}
// the error of the first failing rank becomes the multi-context's error
static int mcollect(phyfoo_mctx* mc, int rc, const char* who) {
    if (!rc) return PHYFOO_OK;
    for (size_t r = 0; r < mc->ctx.size(); ++r)
        if (!mc->ctx[r]->err.empty()) { mc->err = std::string(who) + " (device " + std::to_string(mc->ctx[r]->device) + "): " + mc->ctx[r]->err; return rc; }
    mc->err = std::string(who) + ": failed";
    return rc;
}
static int mreduce(phyfoo_mctx* mc, int r, double* buf, size_t n) {
    const int e = mc->nccl.AllReduce(buf, buf, n, NcclApi::kFloat64, NcclApi::kSum, mc->comm[r], mc->ctx[r]->stream);
    if (e) { mc->ctx[r]->err = std::string("ncclAllReduce: ") + mc->nccl.GetErrorString(e); return PHYFOO_ECUDA; }
    return PHYFOO_OK;
}

extern "C" int phyfoo_init_multi(const int* device_ids, int n_devices, phyfoo_mctx** out) {
    if (!out) return mfail(nullptr, PHYFOO_EINVAL, "phyfoo_init_multi: out is NULL");
    *out = nullptr;
    if (!device_ids || n_devices < 1 || n_devices > 64) return mfail(nullptr, PHYFOO_EINVAL, "phyfoo_init_multi: need 1..64 devices");
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (device_ids[i] == device_ids[j]) return mfail(nullptr, PHYFOO_EINVAL, "phyfoo_init_multi: a device is listed twice");
    phyfoo_mctx* mc = new phyfoo_mctx();
    for (int i = 0; i < n_devices; ++i) {
        phyfoo_ctx* c = nullptr;
        const int rc = phyfoo_init(device_ids[i], &c);
        if (rc) {
            const std::string e = g_init_err;
            for (phyfoo_ctx* p : mc->ctx) phyfoo_destroy(p);
            delete mc;
            return mfail(nullptr, rc, "phyfoo_init_multi: " + e);
        }
        mc->ctx.push_back(c);
    }
    auto bail = [&](int code, const std::string& msg) {
        for (size_t r = 0; r < mc->ctx.size(); ++r) {
            if (r < mc->d_red.size() && mc->d_red[r]) { cudaSetDevice(mc->ctx[r]->device); cudaFree(mc->d_red[r]); }
            phyfoo_destroy(mc->ctx[r]);
        }
        delete mc;
        return mfail(nullptr, code, msg);
    };
    if (n_devices > 1) {
        const std::string e = mc->nccl.load();
        if (!e.empty()) return bail(PHYFOO_EUNSUPPORTED, "phyfoo_init_multi: " + e);
        mc->comm.assign(n_devices, nullptr);
        const int ne = mc->nccl.CommInitAll(mc->comm.data(), n_devices, device_ids);
        if (ne) return bail(PHYFOO_ECUDA, std::string("ncclCommInitAll: ") + mc->nccl.GetErrorString(ne));
    }
    mc->red_cap = 4096;
    mc->d_red.assign(n_devices, nullptr);
    for (int r = 0; r < n_devices; ++r) {
        cudaSetDevice(device_ids[r]);
        if (cudaMalloc(&mc->d_red[r], mc->red_cap * sizeof(double)) != cudaSuccess) return bail(PHYFOO_ENOMEM, "phyfoo_init_multi: cudaMalloc failed");
    }
    mc->workers.start(n_devices);
    *out = mc;
    return PHYFOO_OK;
}

extern "C" void phyfoo_destroy_multi(phyfoo_mctx* mc) {
    if (!mc) return;
    mc->workers.stop();
    for (size_t r = 0; r < mc->ctx.size(); ++r) {
        cudaSetDevice(mc->ctx[r]->device);
        cudaStreamSynchronize(mc->ctx[r]->stream);
        if (r < mc->comm.size() && mc->comm[r]) mc->nccl.CommDestroy(mc->comm[r]);
        if (mc->d_red[r]) cudaFree(mc->d_red[r]);
        phyfoo_destroy(mc->ctx[r]);
    }
    delete mc;
}

extern "C" const char* phyfoo_multi_last_error(const phyfoo_mctx* mc) { return mc ? mc->err.c_str() : g_minit_err.c_str(); }
extern "C" int phyfoo_multi_size(const phyfoo_mctx* mc) { return mc ? (int)mc->ctx.size() : 0; }
extern "C" phyfoo_ctx* phyfoo_multi_ctx(phyfoo_mctx* mc, int rank) { return mc && rank >= 0 && rank < (int)mc->ctx.size() ? mc->ctx[rank] : nullptr; }
extern "C" int phyfoo_multi_set_option(phyfoo_mctx* mc, const char* name, int64_t value) {
    if (!mc) return PHYFOO_EINVAL;
    for (phyfoo_ctx* c : mc->ctx) {
        const int rc = phyfoo_set_option(c, name, value);
        if (rc) return mcollect(mc, rc, "multi_set_option");
    }
    return PHYFOO_OK;
}

// contiguous alignment blocks whose column counts are as equal as the boundaries allow; every device gets at least one
// alignment when there are enough of them (phyfoo_b200/sharding.py::shard_bounds is the same rule)
static std::vector<int32_t> shard_bounds_host(const int64_t* col_off, int n, int world) {
    std::vector<int32_t> cuts(world + 1, 0);
    const int64_t total = col_off[n];
    for (int r = 0; r <= world; ++r) {
        const double target = (double)total * r / world;
        cuts[r] = (int32_t)(std::lower_bound(col_off, col_off + n + 1, target, [](int64_t v, double t) { return (double)v < t; }) - col_off);
    }
    cuts[0] = 0; cuts[world] = n;
    for (int r = 1; r < world; ++r) {
        const int lo = n >= world ? r : std::min(r, n), hi = n >= world ? n - (world - r) : n;
        cuts[r] = std::min(std::max(std::max(cuts[r], cuts[r - 1]), lo), hi);
    }
    return cuts;
}

extern "C" int phyfoo_multi_dataset_create(phyfoo_mctx* mc, int32_t n_aln, int32_t n_species, const int64_t* col_offsets,
                                           const uint8_t* codes, phyfoo_mdataset** out) {
    if (!mc) return PHYFOO_EINVAL;
    if (!out || !col_offsets || (!codes && n_aln > 0)) return mfail(mc, PHYFOO_EINVAL, "multi_dataset_create: NULL argument");
    *out = nullptr;
    const int world = (int)mc->ctx.size();
    if (n_aln < world) return mfail(mc, PHYFOO_EINVAL, "multi_dataset_create: fewer alignments than devices");
    if (col_offsets[0] != 0) return mfail(mc, PHYFOO_EINVAL, "multi_dataset_create: col_offsets[0] must be 0");
    for (int i = 0; i < n_aln; ++i)
        if (col_offsets[i + 1] < col_offsets[i]) return mfail(mc, PHYFOO_EINVAL, "multi_dataset_create: col_offsets must be non-decreasing");
    phyfoo_mdataset* md = new phyfoo_mdataset();
    md->bounds = shard_bounds_host(col_offsets, n_aln, world);
    md->col_off.assign(col_offsets, col_offsets + n_aln + 1);
    md->S = n_species;
    md->ds.assign(world, nullptr);
    const int rc = mc->workers.run([&](int r) {
        const int lo = md->bounds[r], hi = md->bounds[r + 1];
        std::vector<int64_t> off((size_t)(hi - lo) + 1);
        for (int i = lo; i <= hi; ++i) off[i - lo] = col_offsets[i] - col_offsets[lo];
        return phyfoo_dataset_create(mc->ctx[r], hi - lo, n_species, off.data(), codes + (size_t)col_offsets[lo] * n_species, &md->ds[r]);
    });
    if (rc) {
        for (int r = 0; r < world; ++r) phyfoo_dataset_free(mc->ctx[r], md->ds[r]);
        delete md;
        return mcollect(mc, rc, "multi_dataset_create");
    }
    *out = md;
    return PHYFOO_OK;
}
extern "C" int phyfoo_multi_dataset_bounds(const phyfoo_mdataset* md, int32_t* bounds_out) {
    if (!md || !bounds_out) return PHYFOO_EINVAL;
    for (size_t i = 0; i < md->bounds.size(); ++i) bounds_out[i] = md->bounds[i];
    return PHYFOO_OK;
}
extern "C" void phyfoo_multi_dataset_free(phyfoo_mctx* mc, phyfoo_mdataset* md) {
    if (!md) return;
    for (size_t r = 0; r < md->ds.size(); ++r) phyfoo_dataset_free(mc ? mc->ctx[r] : nullptr, md->ds[r]);
    delete md;
}

extern "C" int phyfoo_multi_model_create(phyfoo_mctx* mc, const phyfoo_model_desc* desc, phyfoo_mmodel** out) {
    if (!mc) return PHYFOO_EINVAL;
    if (!out) return mfail(mc, PHYFOO_EINVAL, "multi_model_create: out is NULL");
    *out = nullptr;
    phyfoo_mmodel* mm = new phyfoo_mmodel();
    for (phyfoo_ctx* c : mc->ctx) {
        phyfoo_model* m = nullptr;
        const int rc = phyfoo_model_create(c, desc, &m);
        if (rc) {
            for (size_t r = 0; r < mm->m.size(); ++r) phyfoo_model_free(mc->ctx[r], mm->m[r]);
            delete mm;
            return mcollect(mc, rc, "multi_model_create");
        }
        mm->m.push_back(m);
    }
    *out = mm;
    return PHYFOO_OK;
}
extern "C" int phyfoo_multi_model_set_pi(phyfoo_mctx* mc, phyfoo_mmodel* mm, int32_t position, const double* pi, int32_t n) {
    if (!mc || !mm) return PHYFOO_EINVAL;
    for (size_t r = 0; r < mm->m.size(); ++r) { const int rc = phyfoo_model_set_pi(mc->ctx[r], mm->m[r], position, pi, n); if (rc) return mcollect(mc, rc, "multi_model_set_pi"); }
    return PHYFOO_OK;
}
extern "C" int phyfoo_multi_model_get_pi(phyfoo_mctx* mc, const phyfoo_mmodel* mm, int32_t position, double* pi_out, int32_t n) {
    if (!mc || !mm) return PHYFOO_EINVAL;
    return mcollect(mc, phyfoo_model_get_pi(mc->ctx[0], mm->m[0], position, pi_out, n), "multi_model_get_pi");
}
extern "C" int phyfoo_multi_model_set_branch(phyfoo_mctx* mc, phyfoo_mmodel* mm, int32_t position, int32_t node, double length) {
    if (!mc || !mm) return PHYFOO_EINVAL;
    for (size_t r = 0; r < mm->m.size(); ++r) { const int rc = phyfoo_model_set_branch(mc->ctx[r], mm->m[r], position, node, length); if (rc) return mcollect(mc, rc, "multi_model_set_branch"); }
    return PHYFOO_OK;
}
extern "C" int phyfoo_multi_model_set_mode(phyfoo_mctx* mc, phyfoo_mmodel* mm, int32_t mode) {
    if (!mc || !mm) return PHYFOO_EINVAL;
    for (size_t r = 0; r < mm->m.size(); ++r) { const int rc = phyfoo_model_set_mode(mc->ctx[r], mm->m[r], mode); if (rc) return mcollect(mc, rc, "multi_model_set_mode"); }
    return PHYFOO_OK;
}
extern "C" void phyfoo_multi_model_free(phyfoo_mctx* mc, phyfoo_mmodel* mm) {
    if (!mm) return;
    for (size_t r = 0; r < mm->m.size(); ++r) phyfoo_model_free(mc ? mc->ctx[r] : nullptr, mm->m[r]);
    delete mm;
}

// ZOOPS E-step over all devices: every device scores its block of alignments, [ll, w0, w1] is all-reduced in-stream, the
// per-window / per-alignment outputs land in the caller's arrays at their global positions (alignment order).
extern "C" int phyfoo_multi_zoops_estep(phyfoo_mctx* mc, const phyfoo_mdataset* md, phyfoo_mmodel* fg, phyfoo_mmodel* bg,
                                        const double* logw, const double* strand_logw, const double* aln_weights,
                                        double* gamma_out, double* profile_out, double* w_out, double* ll_out,
                                        int32_t* argmax_off, uint8_t* argmax_strand, phyfoo_stats* stats) {
    if (!mc) return PHYFOO_EINVAL;
    if (!md || !fg || !bg || !logw) return mfail(mc, PHYFOO_EINVAL, "multi_zoops_estep: NULL argument");
    const int world = (int)mc->ctx.size();
    const int W = fg->m[0]->h.W;
    // window offset of every device's block in the global (alignment-major) order
    std::vector<int64_t> woff(world + 1, 0);
    for (int r = 0; r < world; ++r) {
        int64_t n = 0;
        for (int i = md->bounds[r]; i < md->bounds[r + 1]; ++i) n += std::max<int64_t>(0, md->col_off[i + 1] - md->col_off[i] - W + 1);
        woff[r + 1] = woff[r] + n;
    }
    std::vector<int64_t> nw(world, 0);
    std::vector<std::array<unsigned long long, 4>> cnt(world);
    std::vector<float> ms(world, 0.f);
    std::vector<int> launches(world, 0);
    // phase 1: enqueue the kernels; a rank that fails here must not leave the others alone in the collective
    int rc = mc->workers.run([&](int r) {
        phyfoo_ctx* ctx = mc->ctx[r];
        phyfoo_dataset* ds = md->ds[r];
        ctx->err.clear();
        CU(cudaSetDevice(ctx->device));
        const int64_t guess = phyfoo_dataset_windows(ds, W);
        if (gamma_out) CU(ctx->gamma.ensure((size_t)std::max<int64_t>(guess, 1) * sizeof(double)));
        if (profile_out) CU(ctx->profile.ensure((size_t)std::max<int64_t>(guess, 1) * sizeof(double)));
        if (argmax_off) CU(ctx->argoff.ensure((size_t)std::max(ds->n_aln, 1) * sizeof(int32_t)));
        if (argmax_strand) CU(ctx->argstrand.ensure((size_t)std::max(ds->n_aln, 1)));
        const double* d_alnw = nullptr;
        if (aln_weights) {
            CU(ctx->alnw.ensure((size_t)std::max(ds->n_aln, 1) * sizeof(double)));
            CU(cudaMemcpyAsync(ctx->alnw.p, aln_weights + md->bounds[r], (size_t)ds->n_aln * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            d_alnw = ctx->alnw.as<double>();
        }
        ctx->launches = 0; ctx->n_ktimes = 0;
        CU(cudaEventRecord(ctx->ev0, ctx->stream));
        const int e = enqueue_estep(ctx, ds, fg->m[r], bg->m[r], logw, strand_logw, d_alnw, gamma_out ? ctx->gamma.as<double>() : nullptr,
                                    profile_out ? ctx->profile.as<double>() : nullptr, argmax_off ? ctx->argoff.as<int32_t>() : nullptr,
                                    argmax_strand ? ctx->argstrand.as<uint8_t>() : nullptr, mc->d_red[r], &nw[r]);
        if (e) return e;
        ctx->gamma_windows = gamma_out ? nw[r] : -1;
        ctx->gamma_ds = gamma_out ? ds : nullptr; ctx->gamma_width = W;
        return PHYFOO_OK;
    });
    if (rc) return mcollect(mc, rc, "multi_zoops_estep");
    // phase 2: the all-reduce behind the reduction kernel on the same stream, then the copies, then one wait per device
    double r3[3] = {0, 0, 0};
    rc = mc->workers.run([&](int r) {
        phyfoo_ctx* ctx = mc->ctx[r];
        phyfoo_dataset* ds = md->ds[r];
        CU(cudaSetDevice(ctx->device));
        if (world > 1) { const int e = mreduce(mc, r, mc->d_red[r], 3); if (e) return e; }
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        if (r == 0) CU(cudaMemcpyAsync(r3, mc->d_red[0], sizeof(r3), cudaMemcpyDeviceToHost, ctx->stream));
        if (stats) CU(cudaMemcpyAsync(cnt[r].data(), ctx->counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        if (argmax_off) CU(cudaMemcpyAsync(argmax_off + md->bounds[r], ctx->argoff.p, (size_t)ds->n_aln * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (argmax_strand) CU(cudaMemcpyAsync(argmax_strand + md->bounds[r], ctx->argstrand.p, (size_t)ds->n_aln, cudaMemcpyDeviceToHost, ctx->stream));
        if (gamma_out) CU(cudaMemcpyAsync(gamma_out + woff[r], ctx->gamma.p, (size_t)nw[r] * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (profile_out) CU(cudaMemcpyAsync(profile_out + woff[r], ctx->profile.p, (size_t)nw[r] * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaEventElapsedTime(&ms[r], ctx->ev0, ctx->ev1));
        launches[r] = ctx->launches;
        return PHYFOO_OK;
    });
    if (rc) return mcollect(mc, rc, "multi_zoops_estep");
    if (ll_out) *ll_out = r3[0];
    if (w_out) { w_out[0] = r3[1]; w_out[1] = r3[2]; }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        for (int r = 0; r < world; ++r) {
            stats->n_windows += nw[r] * (strand_logw ? 2 : 1); stats->n_columns += md->ds[r]->n_cols;
            stats->kernel_launches += launches[r]; stats->gpu_ms = std::max(stats->gpu_ms, ms[r]);
            stats->sweeps_fg += (int64_t)cnt[r][1]; stats->sweeps_bg += (int64_t)cnt[r][3];
        }
    }
    return PHYFOO_OK;
}

// M-step objective over all devices.  weights == NULL: every device selects on the gamma its last E-step left resident;
// otherwise weights[windows of the whole sample] (alignment-major).  Then the mean fields of the selected windows are fixed
// on every device (phyfoo_fe_prepare).
extern "C" int phyfoo_multi_fe_prepare_selected(phyfoo_mctx* mc, const phyfoo_mdataset* md, phyfoo_mmodel* m, const double* weights,
                                                double t, double q, phyfoo_mfeset** out, int64_t* n_selected) {
    if (!mc) return PHYFOO_EINVAL;
    if (!md || !m || !out) return mfail(mc, PHYFOO_EINVAL, "multi_fe_prepare_selected: NULL argument");
    *out = nullptr;
    const int world = (int)mc->ctx.size();
    const int W = m->m[0]->h.W;
    std::vector<int64_t> woff(world + 1, 0);
    for (int r = 0; r < world; ++r) {
        int64_t n = 0;
        for (int i = md->bounds[r]; i < md->bounds[r + 1]; ++i) n += std::max<int64_t>(0, md->col_off[i + 1] - md->col_off[i] - W + 1);
        woff[r + 1] = woff[r] + n;
    }
    phyfoo_mfeset* mf = new phyfoo_mfeset();
    mf->fe.assign(world, nullptr);
    std::vector<int64_t> nsel(world, 0);
    const int rc = mc->workers.run([&](int r) {
        phyfoo_ctx* ctx = mc->ctx[r];
        ctx->err.clear();
        int64_t cap = std::max<int64_t>(4 * (int64_t)md->ds[r]->n_aln, 1024), n = 0;
        std::vector<int32_t> sa, so; std::vector<double> sw;
        for (;;) {
            sa.resize(cap); so.resize(cap); sw.resize(cap);
            const int e = phyfoo_select(ctx, md->ds[r], W, weights ? weights + woff[r] : nullptr, t, q, sa.data(), so.data(), sw.data(), cap, &n);
            if (e == PHYFOO_ENOMEM && n > cap) { cap = n; continue; }
            if (e) return e;
            break;
        }
        nsel[r] = n;
        return phyfoo_fe_prepare(ctx, md->ds[r], m->m[r], n, sa.data(), so.data(), nullptr, sw.data(), &mf->fe[r]);
    });
    if (rc) {
        for (int r = 0; r < world; ++r) phyfoo_fe_free(mc->ctx[r], mf->fe[r]);
        delete mf;
        return mcollect(mc, rc, "multi_fe_prepare_selected");
    }
    if (n_selected) { *n_selected = 0; for (int64_t v : nsel) *n_selected += v; }
    *out = mf;
    return PHYFOO_OK;
}

// values[0] = f(lambda), values[1 + i] = f(lambda + eps e_i) (eps > 0), summed over the devices in-stream
static int multi_fe_values(phyfoo_mctx* mc, phyfoo_mfeset* mf, int position, const double* lambda, int dim, double eps, std::vector<double>& v) {
    const int world = (int)mc->ctx.size();
    const size_t nv = eps > 0 ? (size_t)dim + 1 : 1;
    if (nv > mc->red_cap) return mfail(mc, PHYFOO_EUNSUPPORTED, "multi_fe: more values than the reduction buffer holds");
    v.assign(nv, 0.0);
    int rc = mc->workers.run([&](int r) {
        mc->ctx[r]->err.clear();
        return phyfoo_fe_values_device(mc->ctx[r], mf->fe[r], position, lambda, dim, eps, mc->d_red[r], mc->ctx[r]->stream);
    });
    if (rc) return mcollect(mc, rc, "multi_fe_values");
    rc = mc->workers.run([&](int r) {
        phyfoo_ctx* ctx = mc->ctx[r];
        CU(cudaSetDevice(ctx->device));
        if (world > 1) { const int e = mreduce(mc, r, mc->d_red[r], nv); if (e) return e; }
        if (r == 0) CU(cudaMemcpyAsync(v.data(), mc->d_red[0], nv * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return PHYFOO_OK;
    });
    return mcollect(mc, rc, "multi_fe_values");
}
extern "C" int phyfoo_multi_fe_eval(phyfoo_mctx* mc, phyfoo_mfeset* mf, int32_t position, const double* lambda, int32_t dim, double* f_out) {
    if (!mc) return PHYFOO_EINVAL;
    if (!mf || !lambda || !f_out) return mfail(mc, PHYFOO_EINVAL, "multi_fe_eval: NULL argument");
    std::vector<double> v;
    const int rc = multi_fe_values(mc, mf, position, lambda, dim, 0.0, v);
    if (rc) return rc;
    *f_out = v[0];
    return PHYFOO_OK;
}
extern "C" int phyfoo_multi_fe_grad(phyfoo_mctx* mc, phyfoo_mfeset* mf, int32_t position, const double* lambda, int32_t dim, double eps,
                                    double* f_out, double* g_out) {
    if (!mc) return PHYFOO_EINVAL;
    if (!mf || !lambda || !g_out) return mfail(mc, PHYFOO_EINVAL, "multi_fe_grad: NULL argument");
    if (!(eps > 0)) return mfail(mc, PHYFOO_EINVAL, "multi_fe_grad: eps must be > 0");
    std::vector<double> v;
    const int rc = multi_fe_values(mc, mf, position, lambda, dim, eps, v);
    if (rc) return rc;
    if (f_out) *f_out = v[0];
    for (int i = 0; i < dim; ++i) g_out[i] = (v[(size_t)i + 1] - v[0]) / eps;     // jstacs NumericalDifferentiableFunction.java:80
    return PHYFOO_OK;
}
extern "C" int64_t phyfoo_multi_fe_size(const phyfoo_mfeset* mf) {
    int64_t n = 0;
    if (mf) for (phyfoo_feset* fe : mf->fe) n += phyfoo_fe_size(fe);
    return n;
}
extern "C" void phyfoo_multi_fe_free(phyfoo_mctx* mc, phyfoo_mfeset* mf) {
    if (!mf) return;
    for (size_t r = 0; r < mf->fe.size(); ++r) phyfoo_fe_free(mc ? mc->ctx[r] : nullptr, mf->fe[r]);
    delete mf;
}

// sufficient statistics over all devices: the expected counts (W * E doubles) and the entropy are all-reduced once; every
// later evaluation is a host-side dot product inside the library (suffstat_host.hpp), without any device call or collective
extern "C" int phyfoo_multi_ss_create(phyfoo_mctx* mc, phyfoo_mfeset* mf, phyfoo_ssobj** out) {
    if (!mc) return PHYFOO_EINVAL;
    if (!mf || !out) return mfail(mc, PHYFOO_EINVAL, "multi_ss_create: NULL argument");
    *out = nullptr;
    const int world = (int)mc->ctx.size();
    const ModelHost& h0 = mf->fe[0]->m->h;
    const size_t WE = (size_t)h0.W * (h0.order == 1 ? h0.E1() : h0.E);
    std::vector<std::vector<double>> part(world, std::vector<double>(WE + 1, 0.0));
    int rc = mc->workers.run([&](int r) { mc->ctx[r]->err.clear(); return phyfoo_fe_suffstats(mc->ctx[r], mf->fe[r], part[r].data(), &part[r][WE]); });
    if (rc) return mcollect(mc, rc, "multi_ss_create");
    if (world > 1) {
        // through the devices and NCCL like every other exchange of this library (W * E + 1 doubles, once per M-step)
        std::vector<double*> d_buf(world, nullptr);
        rc = mc->workers.run([&](int r) {
            phyfoo_ctx* ctx = mc->ctx[r];
            CU(cudaSetDevice(ctx->device));
            CU(ctx->misc.ensure((WE + 1) * sizeof(double)));
            d_buf[r] = ctx->misc.as<double>();
            CU(cudaMemcpyAsync(d_buf[r], part[r].data(), (WE + 1) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            return PHYFOO_OK;
        });
        if (rc) return mcollect(mc, rc, "multi_ss_create");
        rc = mc->workers.run([&](int r) {
            phyfoo_ctx* ctx = mc->ctx[r];
            CU(cudaSetDevice(ctx->device));
            const int e = mreduce(mc, r, d_buf[r], WE + 1);
            if (e) return e;
            if (r == 0) CU(cudaMemcpyAsync(part[0].data(), d_buf[0], (WE + 1) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            return PHYFOO_OK;
        });
        if (rc) return mcollect(mc, rc, "multi_ss_create");
    }
    phyfoo_ssobj* ss = new phyfoo_ssobj();
    ss->o.init(h0, part[0].data(), part[0][WE]);
    *out = ss;
    return PHYFOO_OK;
}
