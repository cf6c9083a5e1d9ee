/* ORACLE (test infrastructure, not product code): CPU restatement of the reference's NN_rhoD
 * forward pass. Follows neural_nets/neural_net_3D.m:47-62 (simulation), :69-73 (mapminmax_apply),
 * :76-78 (tansig_apply = 2/(1+exp(-2n))-1), :81-85 (mapminmax_reverse); the reference calls the
 * MATLAB-Coder build of the same function through ctypes (cut_select_qp.py:579-582).
 * Weights come from the flat blob documented in sdpcutsel-via-nn_b200/nn_weights.py.
 * Sums run k-ascending starting from 0, then the bias is added (repmat(b,1,Q) + W*a), no FMA contraction
 * (compile with -ffp-contract=off). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may call this.
 */
#include <math.h>
#include <stdint.h>

#define MAXH 64

static double nn_eval_one(const double *blob, const double *in)
{
    int n_in = (int)blob[0], L = (int)blob[1], h = (int)blob[2];
    const double *xo = blob + 3, *xg = xo + n_in, *p = xg + n_in;
    double a[MAXH], z[MAXH];
    for (int i = 0; i < n_in; i++) a[i] = (in[i] - xo[i]) * xg[i] + -1.0;
    int cols = n_in;
    for (int l = 0; l < L - 1; l++) {
        const double *W = p, *b = p + h * cols;
        for (int i = 0; i < h; i++) {
            double s = 0.0;
            for (int k = 0; k < cols; k++) s += W[i * cols + k] * a[k];
            z[i] = b[i] + s;
        }
        for (int i = 0; i < h; i++) a[i] = 2.0 / (1.0 + exp(-2.0 * z[i])) - 1.0;
        p = b + h; cols = h;
    }
    double s = 0.0;
    for (int k = 0; k < h; k++) s += p[k] * a[k];
    double y = p[h] + s;
    double y_gain = p[h + 1], y_xoffset = p[h + 2];
    return (y - -1.0) / y_gain + y_xoffset;
}

/* out[i] = NN(in[i*n_in .. ]) for i in [0, m) */
void nn_oracle_eval(const double *blob, const double *in, int64_t m, double *out)
{
    int n_in = (int)blob[0];
    for (int64_t i = 0; i < m; i++) out[i] = nn_eval_one(blob, in + i * n_in);
}

/* Optimality measure of cut_select_qp.py:573-582 for m candidates given gathered slices:
 *   obj = -(sum_k Qs[k]*Xs[k]) * max_elem + NN([x | Qs]) * max_elem       (Qs already divided by max_elem)
 * sum is left-to-right starting from 0 (Python sum(map(mul, ...))). d = subset size, t = d(d+1)/2. */
void opt_measure(const double *blob, int d, const double *xs, const double *Xs, const double *Qs,
                 const double *max_elem, int64_t m, double *out)
{
    int t = d * (d + 1) / 2;
    double in[32];
    for (int64_t i = 0; i < m; i++) {
        double s = 0.0;
        for (int k = 0; k < t; k++) s += Qs[i * t + k] * Xs[i * t + k];
        for (int k = 0; k < d; k++) in[k] = xs[i * d + k];
        for (int k = 0; k < t; k++) in[d + k] = Qs[i * t + k];
        double obj = -s * max_elem[i];
        obj += nn_eval_one(blob, in) * max_elem[i];
        out[i] = obj;
    }
}
