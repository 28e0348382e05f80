/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Plain-C restatement of the CRF decode used as (a) a second, independent
 * checker next to oracle/crf_oracle.py and (b) the decode stage of the CPU baseline in bench.py (the reference's
 * basecaller decode is CUDA-only koi, bonito/crf/basecall.py:36-40, so a CPU arm needs a port of the in-repo definition).
 *
 * Follows SeqdistModel.decode_batch (bonito/crf/model.py:196-199): posteriors (Log semiring forward-backward over the
 * sparse graph of CTC_CRF, bonito/crf/model.py:37-42,47-67) + 1e-8 -> log -> viterbi (bonito/crf/model.py:98-103), with the
 * tie-breaks and the quality definition of oracle/crf_oracle.py::decode_native (lowest edge, lowest final state; phred of the
 * posterior move mass of the emitted base, bonito/util.py:105-112).
 *
 * scores: [N][T][S*4] float32 (no blank column), outputs [N][T] bytes.  Forward-backward in double, log-posteriors and the
 * Viterbi recursion in float32, exactly as the numpy version.  Build: gcc -O2 -fopenmp -shared -fPIC (oracle/build_ref.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static double lse5(const double* x) {
    double m = x[0];
    for (int i = 1; i < 5; ++i) if (x[i] > m) m = x[i];
    double s = 0.0;
    for (int i = 0; i < 5; ++i) s += exp(x[i] - m);
    return m + log(s);
}

static int decode_one(const float* sc, int T, int state_len, float blank, float qscale, float qbias, uint8_t* moves,
                      uint8_t* seq, uint8_t* qual) {
    int S = 1;
    for (int i = 0; i < state_len; ++i) S *= 4;
    const int Q = S / 4;
    double* alpha = (double*)malloc(sizeof(double) * (size_t)(T + 1) * S);
    double* beta = (double*)malloc(sizeof(double) * (size_t)(T + 1) * S);
    float* v = (float*)malloc(sizeof(float) * 2 * S);
    uint8_t* bp = (uint8_t*)malloc((size_t)T * S);
    double* mass = (double*)malloc(sizeof(double) * (size_t)T * 4);
    if (!alpha || !beta || !v || !bp || !mass) return -1;
    /* in-edge e of state s: e = 0 from s (stay, blank), e = 1 + j from j*Q + s/4 (move), score sc[t][s*4 + j] */
    for (int s = 0; s < S; ++s) alpha[s] = 0.0;
    for (int t = 0; t < T; ++t) {
        const float* m = sc + (size_t)t * S * 4;
        const double* a = alpha + (size_t)t * S;
        double* an = alpha + (size_t)(t + 1) * S;
        for (int s = 0; s < S; ++s) {
            double x[5];
            x[0] = (double)blank + a[s];
            for (int j = 0; j < 4; ++j) x[1 + j] = (double)m[s * 4 + j] + a[j * Q + s / 4];
            an[s] = lse5(x);
        }
    }
    for (int s = 0; s < S; ++s) beta[(size_t)T * S + s] = 0.0;
    for (int t = T - 1; t >= 0; --t) {
        const float* m = sc + (size_t)t * S * 4;
        const double* bn = beta + (size_t)(t + 1) * S;
        double* b = beta + (size_t)t * S;
        for (int p = 0; p < S; ++p) {   /* successors of p: itself (stay) and (4p + c) mod S along edge 1 + p/Q */
            double x[5];
            x[0] = (double)blank + bn[p];
            for (int c = 0; c < 4; ++c) {
                const int s = (4 * p + c) % S;
                x[1 + c] = (double)m[s * 4 + p / Q] + bn[s];
            }
            b[p] = lse5(x);
        }
    }
    double logz;
    {
        double mx = alpha[(size_t)T * S];
        for (int s = 1; s < S; ++s) if (alpha[(size_t)T * S + s] > mx) mx = alpha[(size_t)T * S + s];
        double acc = 0.0;
        for (int s = 0; s < S; ++s) acc += exp(alpha[(size_t)T * S + s] - mx);
        logz = mx + log(acc);
    }
    float* vc = v;
    float* vn = v + S;
    for (int s = 0; s < S; ++s) vc[s] = 0.f;
    for (int t = 0; t < T; ++t) {
        const float* m = sc + (size_t)t * S * 4;
        const double* a = alpha + (size_t)t * S;
        const double* bn = beta + (size_t)(t + 1) * S;
        double* ms = mass + (size_t)t * 4;
        ms[0] = ms[1] = ms[2] = ms[3] = 0.0;
        for (int s = 0; s < S; ++s) {
            float best = -INFINITY;
            int arg = 0;
            for (int e = 0; e < 5; ++e) {
                const int prev = e == 0 ? s : (e - 1) * Q + s / 4;
                const double sce = e == 0 ? (double)blank : (double)m[s * 4 + e - 1];
                const double post = exp(a[prev] + sce + bn[s] - logz);
                if (e > 0) ms[s & 3] += post;
                const float lp = logf((float)post + 1e-8f);
                const float cand = lp + vc[prev];
                if (cand > best) { best = cand; arg = e; }
            }
            vn[s] = best;
            bp[(size_t)t * S + s] = (uint8_t)arg;
        }
        float* tmp = vc; vc = vn; vn = tmp;
    }
    int state = 0;
    for (int s = 1; s < S; ++s) if (vc[s] > vc[state]) state = s;
    for (int t = T - 1; t >= 0; --t) {
        const int e = bp[(size_t)t * S + state];
        if (e != 0) {
            const int base = state & 3;
            double err = 1.0 - mass[(size_t)t * 4 + base];
            if (err < 1e-4) err = 1e-4;
            long q = lrint(-10.0 * log10(err) * (double)qscale + (double)qbias) + 33;
            if (q < 33) q = 33;
            if (q > 126) q = 126;
            moves[t] = 1; seq[t] = (uint8_t)"ACGT"[base]; qual[t] = (uint8_t)q;
            state = (e - 1) * Q + state / 4;
        } else {
            moves[t] = 0; seq[t] = 0; qual[t] = 0;
        }
    }
    free(alpha); free(beta); free(v); free(bp); free(mass);
    return 0;
}

int crf_decode_ref(const float* scores, int N, int T, int state_len, float blank, float qscale, float qbias,
                   uint8_t* moves, uint8_t* seq, uint8_t* qual) {
    int S4 = 4, rc = 0;
    for (int i = 0; i < state_len; ++i) S4 *= 4;
#pragma omp parallel for schedule(dynamic)
    for (int n = 0; n < N; ++n) {
        int r = decode_one(scores + (size_t)n * T * S4, T, state_len, blank, qscale, qbias, moves + (size_t)n * T,
                           seq + (size_t)n * T, qual + (size_t)n * T);
        if (r) rc = r;
    }
    return rc;
}
