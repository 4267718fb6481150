// suffstat_host.hpp -- the fixed-q M-step objective in sufficient-statistics form, evaluated on the host INSIDE the library,
// and the position-wise quasi-Newton driver on top of it.
//
// With q fixed (optimizing/AbstractFreeEnergyOptimizer.java:100-102) the objective is linear in the log tables,
//     f(theta) = sum_t sum_e C[t][e] log table_theta,t[e] - ENT,
// with expected counts C from ONE device pass (phyfoo_fe_suffstats; all-reduced over the devices of a multi-context).  Every
// evaluateFunction / evaluateGradientOfFunction of the Jstacs optimizer is then a dot product of 4 + 16 (N - 1) terms -- no
// device call, no collective -- for the stationary distributions (optimizing/FE_byParameter_ForBayesNodeJstacs.java:43-83) and
// for the branch lengths (optimizing/branchlengths/EdgeOptimizer.java:44-74) alike.  A Java caller binds phyfoo_ss_eval /
// phyfoo_ss_grad under its own Optimizer; phyfoo_ss_optimize is the library's stand-in for that optimizer (a BFGS with a
// bracketing line search and the termination rule of optimizing/PreparedConditions.java:20-35; the reference's own
// trajectory is not reproducible), so that a whole M-step is one call.
#pragma once
#include <limits>
#include <stdexcept>
#include <cmath>
#include <algorithm>
#include <functional>
#include <vector>

#include "model_host.hpp"

namespace phyfoo {

struct SuffStatObjective {
    ModelHost h;                       // a private copy of the model: evaluations move ITS parameters
    std::vector<double> counts;        // [W][E]
    double entropy = 0.0;
    std::vector<double> G;             // [W] per-position terms under the current parameters
    long long evaluations = 0;

    int EE = 0;                        // entries per position: E (order 0) or E1 (order-1 network, build_tables1_into)
    // generic mode (backgrounds of order >= 1, netg.cuh): counts over the generic net's table array, tree t owns the slice
    // [g_off[t], g_off[t + 1])
    bool generic = false;
    std::vector<size_t> g_off;
    void init(const ModelHost& model, const double* c, double ent) {
        h = model;
        generic = false;
        EE = h.order == 1 ? h.E1() : h.E;
        counts.assign(c, c + (size_t)h.W * EE);
        entropy = ent;
        G.assign(h.W, 0.0);
        for (int t = 0; t < h.W; ++t) G[t] = term(t, h.pi[t].data());
    }
    void init_generic(const ModelHost& model, const double* c, size_t n_entries, double ent) {
        h = model;
        generic = true;
        g_off.assign(h.W + 1, 0);
        std::vector<double> lt;
        for (int t = 0; t < h.W; ++t) { h.netg_tree_logtab(t, lt); g_off[t + 1] = g_off[t] + lt.size(); }
        if (g_off[h.W] != n_entries) throw std::invalid_argument("suffstat: counts do not match the net's table array");
        counts.assign(c, c + n_entries);
        entropy = ent;
        G.assign(h.W, 0.0);
        for (int t = 0; t < h.W; ++t) G[t] = term(t, h.pi[t].data());
    }
    // sum_e C[t][e] log table[e] for position t under the stationary distribution p (entries with a zero count never enter
    // the sum: their log may be -inf).  Generic mode: p is h.pi[t] (every caller has stored it there already).
    double term(int t, const double* p) const {
        if (generic) {
            std::vector<double> lt;
            h.netg_tree_logtab(t, lt);
            const double* c = &counts[g_off[t]];
            double s = 0.0;
            for (size_t e = 0; e < lt.size(); ++e) if (c[e] != 0.0) s += c[e] * lt[e];
            return s;
        }
        std::vector<double> lt(EE);
        if (h.order == 1) h.build_tables1_into(t, p, lt.data());
        else h.build_tables_into(t, p, lt.data(), nullptr);
        const double* c = &counts[(size_t)t * EE];
        double s = 0.0;
        for (int e = 0; e < EE; ++e) if (c[e] != 0.0) s += c[e] * lt[e];
        return s;
    }
    double value() const {
        double s = 0.0;
        for (double g : G) s += g;
        return s - entropy;
    }
    // evaluateFunction(lambda) of one position: pi_t = softmax(lambda); the object (like the reference's model) is left there
    int dim(int t) const { return generic ? h.ctx_rows(t) * 4 : (h.order == 1 ? h.ctx_rows(t) * 4 : 4); }
    double eval_lambda(int t, const double* lambda) {
        const int d = dim(t);                                  // four parameters per context row, a softmax per row
        std::vector<double> p(d);
        lambda_to_pi(lambda, p.data(), d);
        h.pi[t].assign(p.begin(), p.end());
        G[t] = term(t, p.data());
        ++evaluations;
        return value();
    }
    // branch lengths of one position (or of all: t < 0) from alpha[N] (entry 0 unused)
    double eval_branches(int t, const double* alpha) {
        for (int pos = (t < 0 ? 0 : t); pos < (t < 0 ? h.W : t + 1); ++pos) {
            for (int k = 1; k < h.N; ++k) h.branch[(size_t)pos * h.N + k] = alpha[k];
            G[pos] = term(pos, h.pi[pos].data());
        }
        ++evaluations;
        return value();
    }
};

// ---- jstacs Optimizer.optimize(QUASI_NEWTON_BFGS, ...) restated statement by statement -------------------------------------
// de/jstacs/algorithms/optimization/Optimizer.java: findBracket :122-148, brentsMethod :202-303, quasiNewtonBFGS :1000-1080;
// OneDimensionalFunction.findMin; LimitedMedianStartDistance; termination/*.java.  The same statements in Python:
// phyfoo_b200/optimize.py (tests/test_optimize_host.py holds the two against each other).  Restated from the Java source, not run
// against it (no JVM): given the same start vector and the same function values the iterates follow the reference's.
struct TerminationCondition {         // CombinedCondition: go on while at least `threshold` sub-conditions say so
    int threshold; double small_difference; int iterations; double small_gradient, small_step;   // < 0 = not part of it
    // SmallDifferenceOfFunctionEvaluationsCondition: eps <= |f_last - f|; IterationCondition: iteration < max;
    // SmallGradientConditon: sum |g_i| >= eps; SmallStepCondition: |d|^2 alpha^2 >= eps (the SQUARED step length)
    bool go_on(int it, double f_last, double f_cur, const std::vector<double>& g, const std::vector<double>& d, double alpha) const {
        int votes = 0;
        if (small_difference >= 0) votes += small_difference <= std::fabs(f_last - f_cur);
        if (iterations >= 0) votes += it < iterations;
        if (small_gradient >= 0) { double s = 0; for (double v : g) s += std::fabs(v); votes += s >= small_gradient; }
        if (small_step >= 0) { double s = 0; for (double v : d) s += v * v; votes += s * (alpha * alpha) >= small_step; }
        return votes >= threshold;
    }
};
inline TerminationCondition tc_motif_separated_positions() { return {3, 1e-4, 4, 1e-4, 1e-4}; }   // PreparedConditions.java:20-27

struct BfgsResult { double f; int iterations; long long evaluations; };

namespace jstacs_opt {
constexpr int GUARD = 100000;        // the Java loops have no bound; a function unbounded below would never return
const double LIM_FIB = (std::sqrt(5.0) - 1.0) / 2.0, LIM_FIB_PLUS1 = LIM_FIB + 1.0, LIM_FIB_PLUS2 = LIM_FIB + 2.0;   // Optimizer.java:66-70

// Optimizer.findBracket :122-148: erg = {x_lower, f, x_middle, f, x_upper, f}
template <class F> inline void find_bracket(F&& f, double lower, double f_lower, double start_distance, double (&erg)[6]) {
    erg[0] = lower; erg[1] = f_lower; erg[2] = lower + start_distance; erg[3] = erg[4] = erg[5] = 0.0;
    erg[3] = f(erg[2]);
    if (erg[1] < erg[3]) {
        for (int it = 0; it < GUARD; ++it) {
            erg[4] = erg[2];
            erg[5] = erg[3];
            erg[2] = erg[0] + (1 - LIM_FIB) * (erg[2] - erg[0]);
            erg[3] = f(erg[2]);
            if (!(erg[1] < erg[3])) return;
        }
    } else {
        for (int it = 0; it < GUARD; ++it) {
            erg[4] = LIM_FIB_PLUS2 * erg[2] - LIM_FIB_PLUS1 * erg[0];
            erg[5] = f(erg[4]);
            if (erg[3] >= erg[5]) { erg[0] = erg[2]; erg[1] = erg[3]; erg[2] = erg[4]; erg[3] = erg[5]; }
            if (!(erg[3] == erg[5])) return;
        }
    }
    throw std::runtime_error("findBracket does not terminate on this function (the reference would loop forever)");
}

// Optimizer.brentsMethod :202-303: minimum in [a, b] from an inner point x; returns {x, f(x)}
template <class F> inline void brents_method(F&& f, double a, double x, double fx, double b, double tol, double (&out)[2]) {
    const double c = 1 - LIM_FIB;
    double d = 0, e, eps, xm, p, q, r, tol1, t2, u, v, w, fu, fv, fw, tol3;
    eps = 1.2e-16;
    eps = std::sqrt(eps);
    e = 0.0;
    v = w = x;
    fv = fw = fx;
    tol3 = tol / 3.0;
    xm = .5 * (a + b);
    tol1 = eps * std::fabs(x) + tol3;
    t2 = 2.0 * tol1;
    for (int it = 0; it < GUARD; ++it) {
        if (!(std::fabs(x - xm) > (t2 - .5 * (b - a)))) { out[0] = x; out[1] = fx; return; }
        p = q = r = 0.0;
        if (std::fabs(e) > tol1) {               // fit the parabola
            r = (x - w) * (fx - fv);
            q = (x - v) * (fx - fw);
            p = (x - v) * q - (x - w) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p; else q = -q;
            r = e;
            e = d;
        }
        if ((std::fabs(p) < std::fabs(.5 * q * r)) && (p > q * (a - x)) && (p < q * (b - x))) {
            d = p / q;                           // a parabolic interpolation step
            u = x + d;
            if (((u - a) < t2) || ((b - u) < t2)) {   // f must not be evaluated too close to a or b
                d = tol1;
                if (x >= xm) d = -d;
            }
        } else {                                 // a golden-section step
            if (x < xm) e = b - x; else e = a - x;
            d = c * e;
        }
        if (std::fabs(d) >= tol1) u = x + d;     // f must not be evaluated too close to x
        else if (d > 0.0) u = x + tol1;
        else u = x - tol1;
        fu = f(u);
        if (fx <= fu) { if (u < x) a = u; else b = u; }
        if (fu <= fx) {
            if (u < x) b = x; else a = x;
            v = w; fv = fw; w = x; fw = fx; x = u; fx = fu;
        } else {
            if ((fu <= fw) || (w == x)) { v = w; fv = fw; w = u; fw = fu; }
            else if (!((fu > fv) && (v != x) && (v != w))) { v = u; fv = fu; }
        }
        xm = .5 * (a + b);
        tol1 = eps * std::fabs(x) + tol3;
        t2 = 2.0 * tol1;
    }
    throw std::runtime_error("brentsMethod does not terminate on this function");
}
}  // namespace jstacs_opt

// Optimizer.quasiNewtonBFGS :1000-1080 minimising fun from x (updated in place).  Gradient: forward differences with eps, x[i]
// perturbed in place and restored (NumericalDifferentiableFunction.java:64-84); the function value of an iterate is the one its
// line search found.  start_distance: the initial value of the LimitedMedianStartDistance(10, .) the reference passes.
inline BfgsResult bfgs_minimise(const std::function<double(const double*)>& fun, std::vector<double>& x, const TerminationCondition& tc,
                                double eps = 1e-4, double lin_eps = 1e-3, double start_distance = 0.1) {
    const int n = (int)x.size();
    long long evals = 0;
    auto f_at = [&](const std::vector<double>& z) { ++evals; return fun(z.data()); };
    auto grad = [&](std::vector<double>& z, std::vector<double>& g) {
        const double current = f_at(z);
        for (int i = 0; i < n; ++i) { const double h = z[i]; z[i] += eps; g[i] = (f_at(z) - current) / eps; z[i] = h; }
    };
    // LimitedMedianStartDistance(10, start_distance)
    double memory[10], sorted[10];
    int mem_index = 0;
    for (double& m : memory) m = start_distance;
    std::vector<double> d(n), s(n), v(n), u(n), matrixv(n), g_old(n), g_new(n), tmp(n), matrix((size_t)n * n, 0.0);
    const double inf = std::numeric_limits<double>::infinity();
    double help[2] = {inf, f_at(x)};
    grad(x, g_new);
    for (int c1 = 0; c1 < n; ++c1) { d[c1] = -g_new[c1]; matrix[(size_t)c1 * n + c1] = 1; }
    int i = 0;
    bool next = tc.go_on(i, inf, help[1], g_new, d, help[0]);
    while (next) {
        const double current = help[1];
        std::copy(memory, memory + 10, sorted);
        std::sort(sorted, sorted + 10);
        const double sd = 0.667 * (0.5 * (sorted[4] + sorted[5]));
        if (!(sd > 0)) throw std::runtime_error("The start distance has to be positive.");
        auto along = [&](double a) { for (int c = 0; c < n; ++c) tmp[c] = x[c] + a * d[c]; return f_at(tmp); };   // OneDimensionalSubFunction
        double br[6];
        jstacs_opt::find_bracket(along, 0.0, current, sd, br);
        jstacs_opt::brents_method(along, br[0], br[2], br[3], br[4], lin_eps, help);                                // findMin
        ++i;
        memory[mem_index] = help[0];
        mem_index = (mem_index + 1) % 10;
        for (int c1 = 0; c1 < n; ++c1) x[c1] += d[c1] * help[0];
        g_old = g_new;
        grad(x, g_new);
        next = tc.go_on(i, current, help[1], g_new, d, help[0]);
        if (next) {
            for (int c1 = 0; c1 < n; ++c1) { s[c1] = d[c1] * help[0]; v[c1] = g_new[c1] - g_old[c1]; }
            double vmatrixv = 0, sv = 0, vv = 0;
            for (int c1 = 0; c1 < n; ++c1) {
                sv += s[c1] * v[c1];
                vv += v[c1] * v[c1];
                matrixv[c1] = 0;
                for (int c2 = 0; c2 < n; ++c2) matrixv[c1] += matrix[(size_t)c1 * n + c2] * v[c2];
                vmatrixv += v[c1] * matrixv[c1];
            }
            (void)vv;
            for (int c1 = 0; c1 < n; ++c1) u[c1] = s[c1] / sv - matrixv[c1] / vmatrixv;
            for (int c1 = 0; c1 < n; ++c1) {
                for (int c2 = 0; c2 < n; ++c2)
                    matrix[(size_t)c1 * n + c2] += s[c1] * s[c2] / sv - matrixv[c1] * matrixv[c2] / vmatrixv + vmatrixv * u[c1] * u[c2];
                d[c1] = 0;
                for (int c2 = 0; c2 < n; ++c2) d[c1] += -matrix[(size_t)c1 * n + c2] * g_new[c2];
            }
        }
    }
    return {help[1], i, evals};
}

}  // namespace phyfoo
