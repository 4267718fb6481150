// jit_o0.hpp -- order-0 mean field, kernel compiled at run time for ONE tree (NVRTC, sm_100a).
//
// score_o0f_kernel (tree_pass_f81.cuh) walks a tree "program" held in shared memory: per node and pass it spends more
// instructions on finding its operands (slot numbers, table offsets, leaf lists, the strided state addresses:
// profiles/r02_c3o0_summary.md -- a quarter of the executed instructions are fp64) than on the update itself.  A model's
// tree is fixed for the life of the model, so the pass is generated as straight-line code for that tree: every table entry
// and state word is an immediate offset from one per-thread base, the q of a node lives in registers from its update to its
// last child's, leaf species are literal shift counts, and the only cross-pass state is the children's message sums of the
// nodes that have internal children (4 doubles per such node) plus, when the data set has gaps at all, one scalar per node.
// The arithmetic is meanfield_pass_f81's statement for statement (same helpers, embedded below as NVRTC headers), the kernel
// around the pass is score_o0f_kernel's (window groups, one named barrier per pass, windows fetched ahead).
//
// The library never links NVRTC or the driver API: libnvrtc.so.12 and libcuda.so.1 are opened on first use; when either is
// missing or a compilation fails the caller falls back to score_o0f_kernel (still a CUDA kernel of this library -- there is no
// CPU path) and phyfoo_last_error keeps the reason.
#pragma once
#include <dlfcn.h>

#include <memory>
#include <sstream>

#include "jit_embed.inc"     // JIT_HEADER_NAMES / JIT_HEADER_TEXT: the shared device headers as text (phyfoo_b200/lib.py writes it)

struct JitShape { int T, GL, NG; size_t smem; bool gstate; int state_doubles; };

struct JitKernel {
    void* module = nullptr;      // CUmodule
    void* fn = nullptr;          // CUfunction
    JitShape shape{};
    int regs = 0;
};

// ---- NVRTC / driver entry points, resolved once ----
struct JitApi {
    bool tried = false, ok = false;
    std::string why;
    // nvrtc
    int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(void*, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(void*, size_t*) = nullptr;
    int (*GetCUBIN)(void*, char*) = nullptr;
    int (*GetProgramLogSize)(void*, size_t*) = nullptr;
    int (*GetProgramLog)(void*, char*) = nullptr;
    int (*DestroyProgram)(void**) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    // driver
    int (*ModuleLoadData)(void**, const void*) = nullptr;
    int (*ModuleGetFunction)(void**, void*, const char*) = nullptr;
    int (*ModuleUnload)(void*) = nullptr;
    int (*FuncSetAttribute)(void*, int, int) = nullptr;
    int (*FuncGetAttribute)(int*, int, void*) = nullptr;
    int (*LaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
    bool have_driver = false;
};

static JitApi& jit_api() {
    static JitApi api;
    if (api.tried) return api;
    api.tried = true;
    void* rtc = nullptr;
    for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"})
        if ((rtc = dlopen(name, RTLD_NOW | RTLD_LOCAL))) break;
    if (!rtc) { api.why = "libnvrtc.so.12 not found"; return api; }
#define JIT_SYM(lib, field, sym) *(void**)(&api.field) = dlsym(lib, sym); if (!api.field) { api.why = std::string("missing symbol ") + sym; return api; }
    JIT_SYM(rtc, CreateProgram, "nvrtcCreateProgram");
    JIT_SYM(rtc, CompileProgram, "nvrtcCompileProgram");
    JIT_SYM(rtc, GetCUBINSize, "nvrtcGetCUBINSize");
    JIT_SYM(rtc, GetCUBIN, "nvrtcGetCUBIN");
    JIT_SYM(rtc, GetProgramLogSize, "nvrtcGetProgramLogSize");
    JIT_SYM(rtc, GetProgramLog, "nvrtcGetProgramLog");
    JIT_SYM(rtc, DestroyProgram, "nvrtcDestroyProgram");
    JIT_SYM(rtc, GetErrorString, "nvrtcGetErrorString");
    api.ok = true;
    void* drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!drv) { api.why = "libcuda.so.1 not found (compilation only)"; return api; }
    JIT_SYM(drv, ModuleLoadData, "cuModuleLoadData");
    JIT_SYM(drv, ModuleGetFunction, "cuModuleGetFunction");
    JIT_SYM(drv, ModuleUnload, "cuModuleUnload");
    JIT_SYM(drv, FuncSetAttribute, "cuFuncSetAttribute");
    JIT_SYM(drv, FuncGetAttribute, "cuFuncGetAttribute");
    JIT_SYM(drv, LaunchKernel, "cuLaunchKernel");
#undef JIT_SYM
    api.have_driver = true;
    return api;
}

// ---- the generator ----
// state words of a thread: cs of slot s at 4 s + a, the gap-leaf sum of internal node i at 4 n_slots + i (only with gaps)
static std::string jit_o0_source(const ModelHost& h, const JitShape& sh, bool gaps, int nb) {
    std::ostringstream o;
    const int I = h.I, W = h.W;
    std::vector<int> slot(I, -1);
    int n_slots = 0;
    for (int i = 0; i < I; ++i) if (h.ichild_begin[i + 1] > h.ichild_begin[i]) slot[i] = n_slots++;
    o << "#include \"fp64_math.cuh\"\n#include \"n1n_math.cuh\"\n#include \"tree_pass.cuh\"\n#include \"score_common.cuh\"\n"
         "using namespace phyfoo;\n";
    o << "#define JW " << W << "\n#define JT " << sh.T << "\n#define JGL " << sh.GL << "\n#define JNG " << sh.NG << "\n#define JWPG " << sh.GL / W
      << "\n#define JEF " << h.Ef() << "\n#define JSTATE " << sh.state_doubles << "\n#define JNB " << nb << "\n";
    o << "#define TB(e) tb[(e) * JW]\n";
    if (sh.gstate) o << "#define ST(k) st[(size_t)(k) * gstride]\n";
    else o << "#define ST(k) st[(k) * JT]\n";
    if (gaps)
        o << "// an unobserved leaf below a node whose new q sums to qs (tree_pass_f81.cuh:124-146)\n"
             "__device__ __noinline__ void jit_gap_leaf(double qs, double lpi0, double lpi1, double lpi2, double lpi3, bool update, const double* exptab,\n"
             "                                          N1nAcc& acc, double& tsum) {\n"
             "    const double lpi[4] = {lpi0, lpi1, lpi2, lpi3};\n"
             "    double xl[4], xe[4], el[4];\n"
             "#pragma unroll\n"
             "    for (int a = 0; a < 4; ++a) { xl[a] = qs * lpi[a]; xe[a] = update ? xl[a] : 0.0; }\n"
             "    n1n_exp4(xe, exptab, el);\n"
             "    const double zl = ((el[0] + el[1]) + el[2]) + el[3];\n"
             "    const N1nZR zr = n1n_z_rare(zl);\n"
             "    double t = 0.0, fl = 0.0;\n"
             "#pragma unroll\n"
             "    for (int a = 0; a < 4; ++a) {\n"
             "        const double ql = el[a] * zr.rz;\n"
             "        t = fma(ql, lpi[a], t);\n"
             "        fl = fma(ql, update ? 0.0 : -xl[a], fl);\n"
             "    }\n"
             "    acc.Fq += fl;\n"
             "    acc.Fslow += zr.lz;\n"
             "    tsum += t;\n"
             "}\n";
    o << "__device__ __forceinline__ void jit_pass(const double* __restrict__ tb, double* __restrict__ st, const ColWords cw, const bool update,\n"
         "                                         const double* __restrict__ exptab, N1nAcc& acc" << (sh.gstate ? ", const int gstride" : "") << ") {\n";
    o << "    const double lpi0 = TB(0), lpi1 = TB(1), lpi2 = TB(2), lpi3 = TB(3);\n";
    if (gaps) o << "    const bool col_gap = cw.g != 0;\n";
    for (int i = 0; i < I; ++i) {
        const int p = h.par_internal[i], sl = slot[i], sp = i ? slot[p] : -1;
        const std::string n = std::to_string(i), pn = std::to_string(p);
        const bool has_leaves = h.leaf_begin[i + 1] > h.leaf_begin[i];
        o << "    // internal node " << i << " (tree node " << h.node_of[i] << ")\n";
        o << "    double q" << n << "_0, q" << n << "_1, q" << n << "_2, q" << n << "_3;\n    {\n";
        o << "        double oc[4];\n";
        if (i == 0) o << "        oc[0] = lpi0; oc[1] = lpi1; oc[2] = lpi2; oc[3] = lpi3;\n";
        else {
            const int e0 = 4 + 8 * (h.node_of[i] - 1);
            o << "        const double Lo[4] = {TB(" << e0 << "), TB(" << e0 + 1 << "), TB(" << e0 + 2 << "), TB(" << e0 + 3 << ")};\n";
            o << "        const double D[4] = {TB(" << e0 + 4 << "), TB(" << e0 + 5 << "), TB(" << e0 + 6 << "), TB(" << e0 + 7 << ")};\n";
            o << "        const double qp[4] = {q" << pn << "_0, q" << pn << "_1, q" << pn << "_2, q" << pn << "_3};\n";
            o << "        const double tot = ((qp[0] + qp[1]) + qp[2]) + qp[3];\n";
            o << "#pragma unroll\n        for (int a = 0; a < 4; ++a) oc[a] = fma(qp[a], D[a], tot * Lo[a]);\n";
        }
        if (gaps && has_leaves) o << "        bool any_gap = false;\n";
        for (int k = h.leaf_begin[i]; k < h.leaf_begin[i + 1]; ++k) {
            const int spc = h.leaf_sp[k], e0 = 4 + 8 * (h.leaf_node[k] - 1);
            o << "        {\n";
            if (gaps) o << "            const int x = ((cw.g >> " << spc << ") & 1u) ? 4 : (int)((cw.b >> " << 2 * spc << ") & 3u);\n            if (x < 4) {\n";
            else o << "            const int x = (int)((cw.b >> " << 2 * spc << ") & 3u);\n            {\n";
            o << "                const double lo = tb[" << e0 << " * JW + x * JW], d = tb[" << e0 + 4 << " * JW + x * JW];\n";
            o << "#pragma unroll\n                for (int a = 0; a < 4; ++a) oc[a] += a == x ? lo + d : lo;\n";
            o << "            }" << (gaps ? " else any_gap = true;" : "") << "\n        }\n";
        }
        o << "        double m[4] = {0.0, 0.0, 0.0, 0.0}, ex[4], g[4], e[4];\n";
        if (sl >= 0) o << "        m[0] = ST(" << 4 * sl << "); m[1] = ST(" << 4 * sl + 1 << "); m[2] = ST(" << 4 * sl + 2 << "); m[3] = ST(" << 4 * sl + 3 << ");\n";
        if (gaps) o << "        if (col_gap) { const double tsv = ST(" << 4 * n_slots + i << "); m[0] += tsv; m[1] += tsv; m[2] += tsv; m[3] += tsv; }\n";
        o << "#pragma unroll\n        for (int a = 0; a < 4; ++a) { ex[a] = update ? oc[a] + m[a] : 0.0; g[a] = update ? m[a] : -oc[a]; }\n";
        o << "        n1n_exp4(ex, exptab, e);\n"
             "        const double z = ((e[0] + e[1]) + e[2]) + e[3];\n"
             "        double fu = e[0] * g[0];\n"
             "        fu = fma(e[1], g[1], fu); fu = fma(e[2], g[2], fu); fu = fma(e[3], g[3], fu);\n"
             "        const int zhi = n1n_hi(z);\n"
             "        const unsigned zef = (unsigned)zhi >> 20;\n"
             "        const bool znormal = zef - 24u < 2000u;\n"
             "        double rz = n1n_rcp_normal(z);\n"
             "        if (znormal) { acc.E += (int)zef - 1023; acc.M *= n1n_with_hi(z, (zhi & 0x000fffff) | 0x3ff00000); }\n"
             "        else { const N1nZR zr = n1n_z_rare(z); rz = zr.rz; acc.Fslow += zr.lz; }\n"
             "        acc.Fq = fma(fu, rz, acc.Fq);\n";
        o << "        q" << n << "_0 = e[0] * rz; q" << n << "_1 = e[1] * rz; q" << n << "_2 = e[2] * rz; q" << n << "_3 = e[3] * rz;\n";
        if (sl >= 0) o << "        ST(" << 4 * sl << ") = 0.0; ST(" << 4 * sl + 1 << ") = 0.0; ST(" << 4 * sl + 2 << ") = 0.0; ST(" << 4 * sl + 3 << ") = 0.0;\n";
        if (i != 0) {
            o << "        double sw = q" << n << "_0 * Lo[0];\n";
            o << "        sw = fma(q" << n << "_1, Lo[1], sw); sw = fma(q" << n << "_2, Lo[2], sw); sw = fma(q" << n << "_3, Lo[3], sw);\n";
            for (int l = 0; l < 4; ++l) o << "        ST(" << 4 * sp + l << ") += fma(q" << n << "_" << l << ", D[" << l << "], sw);\n";
        }
        if (gaps) {
            o << "        if (col_gap) {\n            double tsum = 0.0;\n";
            if (has_leaves) {
                o << "            if (any_gap) {\n";
                o << "                const double qs = ((q" << n << "_0 + q" << n << "_1) + q" << n << "_2) + q" << n << "_3;\n";
                for (int k = h.leaf_begin[i]; k < h.leaf_begin[i + 1]; ++k)
                    o << "                if ((cw.g >> " << h.leaf_sp[k] << ") & 1u) jit_gap_leaf(qs, lpi0, lpi1, lpi2, lpi3, update, exptab, acc, tsum);\n";
                o << "            }\n";
            }
            o << "            ST(" << 4 * n_slots + i << ") = tsum;\n        }\n";
        }
        o << "    }\n";
    }
    o << "}\n";
    // the kernel around the pass: score_o0f_kernel with the shapes as constants
    o << R"KERNEL(
extern "C" __global__ void __launch_bounds__(JT, 1) score_o0j_kernel(ScoreFParams p) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    double* s_tab = smem;                                            // [JEF][JW]
    double* s_exp = s_tab + JEF * JW;                                // [64]
    double* s_F = s_exp + 64;                                        // [2][JT]
    long long* s_next = (long long*)(s_F + 2 * JT);                  // [JNG][JWPG]
    double* s_state = (double*)(s_next + JNG * JWPG);
    for (int i = tid; i < JEF * JW; i += JT) s_tab[i] = p.tab[i];
    for (int i = tid; i < 64; i += JT) s_exp[i] = d_phy_exp_tab[i];
    const int grp = tid / JGL, gl = tid - grp * JGL;
    const int ws = gl / JW, t = gl - ws * JW;
    const bool lane_ok = ws < JWPG;
    const int bar_id = 1 + grp;
    long long* my_next = s_next + grp * JWPG + ws;
    const double* tb = s_tab + t;
#if JGSTATE
    double* st = p.gstate + ((size_t)blockIdx.x * JT + tid);
    const int gstride = gridDim.x * JT;
#else
    double* st = s_state + tid;
#endif
    const long long n_items = (long long)p.n_windows * p.n_strands;
    if (lane_ok && t == 0) *my_next = (long long)atomicAdd(p.counter_next, 1ULL);
    __syncthreads();
    bool need = lane_ok, active = lane_ok, upd = false, conv = false, pre_valid = lane_ok;
    int k = 0, buf = 0;
    long long item = -1, pre = lane_ok ? *my_next : 0;
    double old = 0.0;
    __syncthreads();
    ColWords cw; cw.b = 0; cw.g = 0;
    double* const fwin = s_F + grp * JGL + ws * JW;
    for (;;) {
        if (need && active) {
            item = pre;
            pre_valid = false;
            if (t == 0) *my_next = (long long)atomicAdd(p.counter_next, 1ULL);
            if (item >= n_items) active = false;
            else {
                const int si = (int)(item / p.n_windows);
                const int64_t w = item - (long long)si * p.n_windows;
                const int strand = p.n_strands == 2 ? si : p.strand0;
                const int a = find_alignment_hint(p.win_off, p.n_aln, w, p.n_windows);
                const int64_t c0 = p.col_off[a] + (w - p.win_off[a]);
                const int64_t c = strand ? c0 + (JW - 1 - t) : c0 + t;
                cw = load_column(p.bases, p.gaps, p.n_cols, JNB, c);
                if (strand) cw = complement(cw);
                k = 0; conv = false; upd = false; need = false;
            }
        }
        double F = 0.0;
        if (active) {
            N1nAcc acc; acc.Fq = 0.0; acc.M = 1.0; acc.E = 0; acc.Fslow = 0.0;
#if JGSTATE
            jit_pass(tb, st, cw, upd, s_exp, acc, gstride);
#else
            jit_pass(tb, st, cw, upd, s_exp, acc);
#endif
            F = n1n_free_energy(acc);
        }
        s_F[buf * JT + tid] = F;
        if (!named_bar_or(bar_id, JGL, active)) break;
        if (!pre_valid && lane_ok) { pre = *my_next; pre_valid = true; }
        if (active) {
            double Fsum = 0.0;
            const double* fp = fwin + buf * JT;
#pragma unroll
            for (int m = 0; m < JW; ++m) Fsum += fp[m];
            bool done;
            if (!upd) { old = Fsum; upd = true; done = false; }
            else {
                if (fabs(old - Fsum) < p.tol) conv = true;
                old = Fsum; ++k;
                done = !((k < p.max_sweeps && !conv) || k < p.min_sweeps);
            }
            if (done) {
                if (t == 0) {
                    p.out[item] = -Fsum;
                    if (p.sweeps) p.sweeps[item] = k;
                    atomicAdd(p.counter_sweeps, (unsigned long long)k);
                }
                need = true;
            }
        }
        buf ^= 1;
    }
}
)KERNEL";
    std::string src = o.str();
    const std::string def = std::string("#define JGSTATE ") + (sh.gstate ? "1" : "0") + "\n";
    return def + src;
}

static size_t jit_o0_fixed_smem(const ModelHost& h, int T, int GL, int NG) {
    return ((size_t)h.Ef() * h.W + 64 + 2 * (size_t)T) * sizeof(double) + (size_t)NG * (GL / h.W) * sizeof(long long) + 64;
}

// block shape: groups of lcm-ish lanes as in launch_score_f81; as many groups as shared memory (state) and 1024 threads allow
static bool jit_o0_shape(const ModelHost& h, bool gaps, size_t budget, int max_threads, JitShape* out) {
    const int W = h.W;
    int best_nwg = 0; double best_use = 0.0;
    for (int nwg = 1; nwg <= 8; ++nwg) {
        if (32 * nwg < W) continue;
        const double use = (double)((32 * nwg) / W * W) / (32 * nwg);
        if (use > best_use + 1e-9) { best_use = use; best_nwg = nwg; }
    }
    if (!best_nwg) return false;
    const int GL = 32 * best_nwg;
    int n_slots = 0;
    for (int i = 0; i < h.I; ++i) n_slots += h.ichild_begin[i + 1] > h.ichild_begin[i];
    const int state_doubles = 4 * n_slots + (gaps ? h.I : 0);
    const size_t per_thread = (size_t)state_doubles * sizeof(double);
    const int max_groups = std::max(1, std::min(15, max_threads / GL));
    JitShape s{};
    s.GL = GL; s.state_doubles = state_doubles;
    for (int ng = max_groups; ng >= 1; --ng) {
        const int T = ng * GL;
        if (jit_o0_fixed_smem(h, T, GL, ng) + per_thread * T <= budget) {
            if (T >= 256 || ng == max_groups) { s.NG = ng; s.T = T; s.gstate = false; s.smem = jit_o0_fixed_smem(h, T, GL, ng) + per_thread * T; *out = s; return true; }
            break;
        }
    }
    s.NG = max_groups; s.T = max_groups * GL; s.gstate = true; s.smem = jit_o0_fixed_smem(h, s.T, GL, s.NG);
    if (s.smem > budget) return false;
    *out = s;
    return true;
}

// compile (and, with a driver, load) the kernel of `src`; cubin_size_out for the compile-only check
static int jit_compile(const std::string& src, std::string& log, std::vector<char>* cubin) {
    JitApi& api = jit_api();
    if (!api.ok) { log = api.why; return -1; }
    void* prog = nullptr;
    int rc = api.CreateProgram(&prog, src.c_str(), "score_o0j.cu", JIT_N_HEADERS, JIT_HEADER_TEXT, JIT_HEADER_NAMES);
    if (rc) { log = std::string("nvrtcCreateProgram: ") + api.GetErrorString(rc); return -1; }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
    rc = api.CompileProgram(prog, 3, opts);
    size_t n = 0;
    if (api.GetProgramLogSize(prog, &n) == 0 && n > 1) { log.resize(n); api.GetProgramLog(prog, &log[0]); }
    if (rc) { log = std::string("nvrtcCompileProgram: ") + api.GetErrorString(rc) + "\n" + log; api.DestroyProgram(&prog); return -1; }
    size_t sz = 0;
    if (api.GetCUBINSize(prog, &sz) || sz == 0) { log = "nvrtcGetCUBINSize failed"; api.DestroyProgram(&prog); return -1; }
    cubin->resize(sz);
    api.GetCUBIN(prog, cubin->data());
    api.DestroyProgram(&prog);
    return 0;
}
