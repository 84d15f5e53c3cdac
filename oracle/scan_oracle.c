/*
 * oracle/scan_oracle.c — TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * CPU restatement of the reference selective scan:
 *   forward : selective_scan_ref, kernels/selective_scan/test_selective_scan.py:168-234
 *             (same arithmetic as selective_scan_torch, basicsr/vmamba/models/csms6s.py:29-72):
 *               delta = delta + delta_bias; delta = softplus(delta)       (:186-189, F.softplus threshold 20)
 *               deltaA = exp(delta * A); deltaB_u = delta * B * u         (:206-213)
 *               x = deltaA * x + deltaB_u; y = sum_n x * C                (:218-225)
 *               out = y + u * D                                           (:232)
 *   backward: what autograd derives from that loop; written out analytically, following the recurrences
 *             the reference CUDA kernel uses (cusoflex/selective_scan_bwd_kernel_oflex.cuh:205-223):
 *               g_t = C_t dout_t + a_{t+1} g_{t+1};  du += g d B; ddelta += g u B + g A (h_t - b_t);
 *               dA += g d (h_t - b_t); dB += g d u; dC += dout h; dD += dout u; softplus' = sigmoid(raw)
 *
 * The forward runs in fp32 in the reference's operation order (sequential in t), so it can be pinned
 * against the reference to ~1 ulp; a second entry point runs everything in fp64 ("exact" oracle) and is
 * what error budgets of the parallel GPU scan are measured against.
 * Pinned against the real reference by tests/golden/make_golden.py -> tests/golden/scan_*.npz.
 *
 * Layout (all contiguous, fp32): u, delta, out, dout: (Bt, KD, L); A: (KD, N); B, C: (Bt, G, N, L);
 * D, delta_bias: (KD) or NULL; last_state: (Bt, KD, N) or NULL.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
static inline double softplus_d(double x) { return x > 20.0 ? x : log1p(exp(x)); }

/* fp32 forward in reference order */
void oracle_scan_fwd_f32(const float* u, const float* delta, const float* A, const float* B, const float* C,
                         const float* D, const float* delta_bias, int delta_softplus,
                         int Bt, int KD, int L, int N, int G,
                         float* out, float* last_state)
{
    const int Dg = KD / G;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < Bt; ++b) {
        for (int d = 0; d < KD; ++d) {
            const int g = d / Dg;
            const float* ur = u + ((int64_t)b * KD + d) * L;
            const float* dr = delta + ((int64_t)b * KD + d) * L;
            float* outr = out + ((int64_t)b * KD + d) * L;
            const float bias = delta_bias ? delta_bias[d] : 0.f;
            const float Dv = D ? D[d] : 0.f;
            float* h = (float*)calloc((size_t)N, sizeof(float));
            for (int t = 0; t < L; ++t) {
                float dl = dr[t];
                if (delta_bias) dl = dl + bias;
                if (delta_softplus) dl = softplus_f(dl);
                float y = 0.f;
                for (int n = 0; n < N; ++n) {
                    const float Bv = B[(((int64_t)b * G + g) * N + n) * L + t];
                    const float Cv = C[(((int64_t)b * G + g) * N + n) * L + t];
                    const float dA = expf(dl * A[(int64_t)d * N + n]);
                    const float dBu = dl * Bv * ur[t];
                    h[n] = dA * h[n] + dBu;
                    y += h[n] * Cv;
                }
                outr[t] = D ? y + ur[t] * Dv : y;
            }
            if (last_state) memcpy(last_state + ((int64_t)b * KD + d) * N, h, (size_t)N * sizeof(float));
            free(h);
        }
    }
}

/* fp64 forward + backward. Any of the gradient pointers may be NULL (dout == NULL -> forward only).
 * dB/dC are (Bt, G, N, L) and are summed over the KD/G rows of a group; dA:(KD,N); dD, ddelta_bias:(KD). */
void oracle_scan_f64(const float* u, const float* delta, const float* A, const float* B, const float* C,
                     const float* D, const float* delta_bias, int delta_softplus,
                     int Bt, int KD, int L, int N, int G,
                     double* out, double* last_state,
                     const float* dout,
                     double* du, double* ddelta, double* dA, double* dB, double* dC, double* dD, double* ddelta_bias)
{
    const int Dg = KD / G;
    const int do_bwd = dout != NULL;
    if (do_bwd) {
        memset(dA, 0, sizeof(double) * (size_t)KD * N);
        memset(dB, 0, sizeof(double) * (size_t)Bt * G * N * L);
        memset(dC, 0, sizeof(double) * (size_t)Bt * G * N * L);
        if (dD) memset(dD, 0, sizeof(double) * (size_t)KD);
        if (ddelta_bias) memset(ddelta_bias, 0, sizeof(double) * (size_t)KD);
    }
    /* parallel over (b, g): rows of a group accumulate into the same dB/dC slab, so they stay in one thread;
     * dA/dD/ddelta_bias are per-row and summed over b with a critical section. */
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int b = 0; b < Bt; ++b) {
        for (int g = 0; g < G; ++g) {
            double* hs = (double*)malloc(sizeof(double) * (size_t)L * N);   /* h_t per (t, n) */
            double* dls = (double*)malloc(sizeof(double) * (size_t)L);
            double* gn = (double*)malloc(sizeof(double) * (size_t)N);
            for (int dd = 0; dd < Dg; ++dd) {
                const int d = g * Dg + dd;
                const float* ur = u + ((int64_t)b * KD + d) * L;
                const float* dr = delta + ((int64_t)b * KD + d) * L;
                const double bias = delta_bias ? (double)delta_bias[d] : 0.0;
                const double Dv = D ? (double)D[d] : 0.0;
                for (int n = 0; n < N; ++n) gn[n] = 0.0;
                /* forward */
                for (int t = 0; t < L; ++t) {
                    double raw = (double)dr[t] + bias;
                    double dl = delta_softplus ? softplus_d(raw) : raw;
                    dls[t] = dl;
                    double y = 0.0;
                    for (int n = 0; n < N; ++n) {
                        const int64_t bc = (((int64_t)b * G + g) * N + n) * L + t;
                        const double a = exp(dl * (double)A[(int64_t)d * N + n]);
                        const double bb = dl * (double)B[bc] * (double)ur[t];
                        gn[n] = a * gn[n] + bb;
                        hs[(int64_t)t * N + n] = gn[n];
                        y += gn[n] * (double)C[bc];
                    }
                    if (out) out[((int64_t)b * KD + d) * L + t] = y + (double)ur[t] * Dv;
                }
                if (last_state) for (int n = 0; n < N; ++n) last_state[((int64_t)b * KD + d) * N + n] = gn[n];
                if (!do_bwd) continue;
                /* backward */
                const float* gor = dout + ((int64_t)b * KD + d) * L;
                for (int n = 0; n < N; ++n) gn[n] = 0.0;   /* now: a_{t+1} * g_{t+1} */
                double dD_acc = 0.0, dbias_acc = 0.0;
                double* dA_row = (double*)calloc((size_t)N, sizeof(double));
                for (int t = L - 1; t >= 0; --t) {
                    const double dl = dls[t];
                    const double uu = (double)ur[t];
                    const double go = (double)gor[t];
                    double du_t = Dv * go;
                    double ddl = 0.0;
                    dD_acc += go * uu;
                    for (int n = 0; n < N; ++n) {
                        const int64_t bc = (((int64_t)b * G + g) * N + n) * L + t;
                        const double An = (double)A[(int64_t)d * N + n];
                        const double a = exp(dl * An);
                        const double Bv = (double)B[bc], Cv = (double)C[bc];
                        const double h = hs[(int64_t)t * N + n];
                        const double hprev = t > 0 ? hs[(int64_t)(t - 1) * N + n] : 0.0;
                        const double gt = Cv * go + gn[n];
                        du_t += gt * dl * Bv;
                        ddl += gt * uu * Bv + gt * An * a * hprev;
                        dA_row[n] += gt * dl * a * hprev;
                        dB[bc] += gt * dl * uu;
                        dC[bc] += go * h;
                        gn[n] = a * gt;
                    }
                    if (delta_softplus) {
                        const double raw = (double)dr[t] + bias;
                        if (raw <= 20.0) ddl = ddl / (1.0 + exp(-raw));
                    }
                    du[((int64_t)b * KD + d) * L + t] = du_t;
                    ddelta[((int64_t)b * KD + d) * L + t] = ddl;
                    dbias_acc += ddl;
                }
#pragma omp critical
                {
                    for (int n = 0; n < N; ++n) dA[(int64_t)d * N + n] += dA_row[n];
                    if (dD) dD[d] += dD_acc;
                    if (ddelta_bias) ddelta_bias[d] += dbias_acc;
                }
                free(dA_row);
            }
            free(hs); free(dls); free(gn);
        }
    }
}
