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
    void init(const ModelHost& model, const double* c, double ent) {
        h = model;
        EE = h.order == 1 ? h.E1() : h.E;
        counts.assign(c, c + (size_t)h.W * EE);
        entropy = ent;
        G.assign(h.W, 0.0);
        for (int t = 0; t < h.W; ++t) G[t] = term(t, h.pi[t].data());
    }
    // sum_e C[t][e] log table[e] for position t under the stationary distribution p (entries with a zero count never enter
    // the sum: their log may be -inf)
    double term(int t, const double* p) const {
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
    int dim(int t) const { return h.order == 1 ? h.ctx_rows(t) * 4 : 4; }
    double eval_lambda(int t, const double* lambda) {
        double p[16];
        const int d = dim(t);                                  // an order-1 position t >= 1: four context rows, a softmax per row
        lambda_to_pi(lambda, p, d);
        h.pi[t].assign(p, p + d);
        G[t] = term(t, p);
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

// ---- stand-in for jstacs Optimizer.optimize(QUASI_NEWTON_BFGS, ...) (Optimizer.java:1000-1080): minimises fun ----
struct TerminationCondition {         // jstacs CombinedCondition: go on while at least `threshold` sub-conditions say so
    int threshold; double small_difference; int iterations; double small_gradient, small_step;   // < 0 = not part of it
    bool go_on(int it, double f_last, double f_cur, const std::vector<double>& g, double step) const {
        int votes = 0;
        if (small_difference >= 0) votes += std::fabs(f_last - f_cur) >= small_difference;
        if (iterations >= 0) votes += it < iterations;
        if (small_gradient >= 0) { double s = 0; for (double v : g) s += std::fabs(v); votes += s >= small_gradient; }
        if (small_step >= 0) votes += step >= small_step;
        return votes >= threshold;
    }
};
inline TerminationCondition tc_motif_separated_positions() { return {3, 1e-4, 4, 1e-4, 1e-4}; }   // PreparedConditions.java:20-27

struct BfgsResult { double f; int iterations; long long evaluations; };

// fun(x) -> value.  Gradient: forward differences with eps, x[i] perturbed in place and restored, a closing evaluation at x
// (jstacs NumericalDifferentiableFunction.java:64-84)
inline BfgsResult bfgs_minimise(const std::function<double(const double*)>& fun, std::vector<double>& x, const TerminationCondition& tc,
                                double eps = 1e-4, double lin_eps = 1e-3, double start_distance = 0.1) {
    const int n = (int)x.size();
    long long evals = 0;
    auto f_at = [&](const std::vector<double>& z) { ++evals; return fun(z.data()); };
    auto grad = [&](std::vector<double>& z, std::vector<double>& g) {
        const double f0 = f_at(z);
        for (int i = 0; i < n; ++i) { const double keep = z[i]; z[i] += eps; g[i] = (f_at(z) - f0) / eps; z[i] = keep; }
        f_at(z);
        return f0;
    };
    std::vector<double> H((size_t)n * n, 0.0), g(n), d(n), xn(n), gn(n), s(n), y(n), tmp(n), steps;
    auto identity = [&]() { std::fill(H.begin(), H.end(), 0.0); for (int i = 0; i < n; ++i) H[(size_t)i * n + i] = 1.0; };
    identity();
    double f = grad(x, g);
    int it = 0;
    const double gr = 1.618033988749895;
    for (;;) {
        double gd = 0.0;
        for (int i = 0; i < n; ++i) { double v = 0; for (int j = 0; j < n; ++j) v -= H[(size_t)i * n + j] * g[j]; d[i] = v; gd += g[i] * v; }
        if (gd >= 0) { identity(); for (int i = 0; i < n; ++i) d[i] = -g[i]; }
        double norm = 0.0;
        for (double v : d) norm += v * v;
        norm = std::sqrt(norm);
        if (norm == 0.0) break;
        double start = start_distance;
        if (!steps.empty()) {                                    // LimitedMedianStartDistance(10, 0.1): 0.667 x median of the last 10
            std::vector<double> last(steps.end() - std::min<size_t>(10, steps.size()), steps.end());
            std::sort(last.begin(), last.end());
            const size_t m = last.size();
            start = 0.667 * (m % 2 ? last[m / 2] : 0.5 * (last[m / 2 - 1] + last[m / 2]));
        }
        start = std::max(start / norm, 1e-8);
        auto along = [&](double a) { for (int i = 0; i < n; ++i) tmp[i] = x[i] + a * d[i]; return f_at(tmp); };
        // bracket by golden-ratio expansion / contraction, then golden section to lin_eps
        double alpha = 0.0, f_new = f;
        {
            double a = 0.0, fa = f, b = start, fb = along(b), c, fc;
            bool ok = true;
            if (fb < fa) {
                c = b * (1 + gr); fc = along(c);
                while (fc < fb && c < 1e6) { a = b; fa = fb; b = c; fb = fc; c = b * (1 + gr); fc = along(c); }
            } else {
                c = b; fc = fb; b = c / (1 + gr); fb = along(b);
                while (fb >= fa && b > 1e-12) { c = b; fc = fb; b = c / (1 + gr); fb = along(b); }
                if (fb >= fa) ok = false;
            }
            if (ok) {
                while ((c - a) > lin_eps * std::max(1.0, std::fabs(b))) {
                    if ((c - b) > (b - a)) {
                        const double t = b + (c - b) / (1 + gr), ft = along(t);
                        if (ft < fb) { a = b; fa = fb; b = t; fb = ft; } else { c = t; fc = ft; }
                    } else {
                        const double t = b - (b - a) / (1 + gr), ft = along(t);
                        if (ft < fb) { c = b; fc = fb; b = t; fb = ft; } else { a = t; fa = ft; }
                    }
                }
                alpha = b; f_new = fb;
            }
            (void)fc; (void)f_new;
        }
        double sy = 0.0, step = 0.0;
        for (int i = 0; i < n; ++i) { s[i] = alpha * d[i]; xn[i] = x[i] + s[i]; step += s[i] * s[i]; }
        step = std::sqrt(step);
        const double f2 = grad(xn, gn);
        for (int i = 0; i < n; ++i) { y[i] = gn[i] - g[i]; sy += s[i] * y[i]; }
        if (sy > 1e-14) {                                        // H <- (I - rho s y^T) H (I - rho y s^T) + rho s s^T
            const double rho = 1.0 / sy;
            std::vector<double> Hy(n, 0.0);
            for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Hy[i] += H[(size_t)i * n + j] * y[j];
            double yHy = 0.0;
            for (int i = 0; i < n; ++i) yHy += y[i] * Hy[i];
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j)
                    H[(size_t)i * n + j] += -rho * (s[i] * Hy[j] + Hy[i] * s[j]) + (rho * rho * yHy + rho) * s[i] * s[j];
        }
        steps.push_back(step);
        ++it;
        const bool go = tc.go_on(it, f, f2, gn, step);
        x = xn; f = f2; g = gn;
        if (!go || alpha == 0.0) break;
    }
    return {f, it, evals};
}

}  // namespace phyfoo
