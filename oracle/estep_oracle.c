/* estep_oracle.c -- CPU restatement (float64, pthreads) of the CPD E-step.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
 * CPU-baseline leg may load the shared object built from this file (oracle/_build/).
 *
 * Follows probreg/cpd.py:71-88 of neka-nat/probreg v0.3.7 column by column:
 *   :74-76  K_mn = exp(-|t_source_m - target_n|^2 / (2 sigma2))
 *   :78-79  c = (2 pi sigma2)^(D/2) * w/(1-w) * M/N           (N = n_global)
 *   :80-82  den_n = sum_m K_mn; den_n == 0 -> float32 eps; den_n += c
 *   :84-87  P = K/den; pt1_n = sum_m P_mn; p1_m = sum_n P_mn; px_m = sum_n P_mn target_n
 * Same arithmetic as oracle/cpd_oracle.py (which is pinned against the reference's own
 * outputs); it exists because the numpy version needs ~40 ns per pair and one core, which
 * makes 20k x 20k parity checks and the 100k-point CPU baseline impractical.
 * tests/test_oracle.py checks this file against the numpy oracle.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static int g_threads = 0;
int oracle_num_threads(void) {
    if (g_threads <= 0) {
        long nc = sysconf(_SC_NPROCESSORS_ONLN);
        g_threads = nc < 1 ? 1 : (nc > 64 ? 64 : (int)nc);
    }
    return g_threads;
}
void oracle_set_num_threads(int t) { g_threads = t < 1 ? 1 : t; }

typedef struct {
    const double *t_source, *target;
    long m, j0, j1;
    int dim;
    double inv, c;
    double *pt1, *lp1, *lpx;   /* lp1/lpx: this worker's private accumulators */
    int ok;
} job_t;

static void* worker(void* arg) {
    job_t* jb = (job_t*)arg;
    const double eps32 = 1.1920928955078125e-07;
    const long m = jb->m;
    const int dim = jb->dim;
    double* k = (double*)malloc(sizeof(double) * (size_t)m);
    if (!k) { jb->ok = 0; return 0; }
    for (long j = jb->j0; j < jb->j1; ++j) {
        const double* x = jb->target + j * dim;
        double den = 0.0;                                   /* cpd.py:80 */
        for (long i = 0; i < m; ++i) {
            double d2 = 0.0;
            for (int a = 0; a < dim; ++a) {
                const double d = jb->t_source[i * dim + a] - x[a];
                d2 += d * d;
            }
            k[i] = exp(-d2 * jb->inv);                      /* cpd.py:74-76 */
            den += k[i];
        }
        if (den == 0.0) den = eps32;                        /* cpd.py:81 */
        den += jb->c;                                       /* cpd.py:82 */
        double col = 0.0;
        for (long i = 0; i < m; ++i) {
            const double p = k[i] / den;                    /* cpd.py:84 */
            col += p;
            jb->lp1[i] += p;                                /* cpd.py:86 */
            for (int a = 0; a < dim; ++a) jb->lpx[i * dim + a] += p * x[a];   /* cpd.py:87 */
        }
        jb->pt1[j] = col;                                   /* cpd.py:85 */
    }
    free(k);
    jb->ok = 1;
    return 0;
}

/* t_source: m x dim, target: n x dim (row-major doubles). pt1: n, p1: m, px: m x dim. */
int oracle_estep(const double* t_source, long m, const double* target, long n, int dim, double sigma2, double w,
                 long n_global, double* pt1, double* p1, double* px, double* n_p) {
    int nt = oracle_num_threads();
    if (nt > n) nt = (int)n;
    double c = pow(2.0 * M_PI * sigma2, dim * 0.5);         /* cpd.py:78 */
    c *= w / (1.0 - w) * (double)m / (double)n_global;      /* cpd.py:79 */
    job_t* jobs = (job_t*)calloc((size_t)nt, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)nt, sizeof(pthread_t));
    double* acc = (double*)calloc((size_t)nt * (size_t)m * (dim + 1), sizeof(double));
    if (!jobs || !th || !acc) { free(jobs); free(th); free(acc); return -1; }
    for (int t = 0; t < nt; ++t) {
        job_t* jb = &jobs[t];
        jb->t_source = t_source; jb->target = target; jb->m = m; jb->dim = dim;
        jb->j0 = n * t / nt; jb->j1 = n * (t + 1) / nt;
        jb->inv = 1.0 / (2.0 * sigma2); jb->c = c; jb->pt1 = pt1;
        jb->lp1 = acc + (size_t)t * m * (dim + 1);
        jb->lpx = jb->lp1 + m;
        pthread_create(&th[t], 0, worker, jb);
    }
    int ok = 1;
    for (int t = 0; t < nt; ++t) { pthread_join(th[t], 0); ok &= jobs[t].ok; }
    memset(p1, 0, sizeof(double) * (size_t)m);
    memset(px, 0, sizeof(double) * (size_t)m * dim);
    for (int t = 0; t < nt; ++t) {
        for (long i = 0; i < m; ++i) p1[i] += jobs[t].lp1[i];
        for (long i = 0; i < m * dim; ++i) px[i] += jobs[t].lpx[i];
    }
    double s = 0.0;
    for (long i = 0; i < m; ++i) s += p1[i];                /* cpd.py:88 */
    *n_p = s;
    free(jobs); free(th); free(acc);
    return ok ? 0 : -1;
}
