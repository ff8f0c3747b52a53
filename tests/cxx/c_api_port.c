/* tests/cxx/c_api_port.c -- the reference's C-interface check (tests/unit/c-api.c:45-170 driven by
 * tests/unit/test-c-api.cc:40-108) restated for the B200 library: a plain C program that owns a
 * Libint_t, fills the per-primitive prerequisites itself, calls
 * libint2_build_eri[am1][am2][am3][am4](Libint_t*) and reads targets[0].  Compiled as C against the
 * reference's own <libint2.h> plus the generated headers of include/libint2/util/generated, linked
 * against liblibint_b200_iface.so.  Prints every integral with 17 significant digits; the Python test
 * (tests/test_gpu_iface.py) compares them with the reference Engine (oracle) times the shells'
 * normalization factors, as test-c-api.cc does.
 *
 *   c_api_port                 one contracted quartet per canonical class l <= 2, fixed geometry
 */
#include <libint2.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795
#endif

#if !(defined(LIBINT2_SUPPORT_ERI) && LIBINT2_MAX_AM_eri >= 1)
#error "the library behind the headers must provide the eri task"
#endif

/* F_m(T), m = 0..mmax: series for the top order, downward recursion (the C test links its own
 * calc_f too, tests/unit/c-api.c:36) */
static void boys(double* F, double T, int mmax) {
  const double eT = exp(-T);
  if (T > 35.0) {
    int m;
    F[0] = 0.5 * sqrt(M_PI / T) * erf(sqrt(T));
    for (m = 0; m < mmax; ++m) F[m + 1] = ((2 * m + 1) * F[m] - eT) / (2 * T);
    return;
  }
  {
    double term = 1.0 / (2 * mmax + 1), sum = term;
    int k, m;
    for (k = 1; k < 400; ++k) {
      term *= 2 * T / (2 * mmax + 2 * k + 1);
      sum += term;
      if (term < 1e-17 * sum) break;
    }
    F[mmax] = eT * sum;
    for (m = mmax; m > 0; --m) F[m - 1] = (2 * T * F[m] + eT) / (2 * m - 1);
  }
}

typedef struct {
  int l, K;
  double alpha[3], coef[3], O[3];
} shell_t;

#define SET_F(p, m, v) (&(p)->LIBINT_T_SS_EREP_SS(0)[0])[m] = (v)

/* prerequisites of one primitive quartet, as include/libint2/engine.impl.h:1331-1641 defines them */
static void fill_primitive(Libint_t* p, const shell_t* s, const int* ip, int L) {
  const double a1 = s[0].alpha[ip[0]], a2 = s[1].alpha[ip[1]], a3 = s[2].alpha[ip[2]], a4 = s[3].alpha[ip[3]];
  const double c = s[0].coef[ip[0]] * s[1].coef[ip[1]] * s[2].coef[ip[2]] * s[3].coef[ip[3]];
  const double *A = s[0].O, *B = s[1].O, *C = s[2].O, *D = s[3].O;
  const double gp = a1 + a2, gq = a3 + a4, gpq = gp + gq, rho = gp * gq / gpq;
  double P[3], Q[3], W[3], AB2 = 0, CD2 = 0, PQ2 = 0, F[32], pfac;
  int k, m;
  for (k = 0; k < 3; ++k) {
    P[k] = (a1 * A[k] + a2 * B[k]) / gp;
    Q[k] = (a3 * C[k] + a4 * D[k]) / gq;
    W[k] = (gp * P[k] + gq * Q[k]) / gpq;
    AB2 += (A[k] - B[k]) * (A[k] - B[k]);
    CD2 += (C[k] - D[k]) * (C[k] - D[k]);
    PQ2 += (P[k] - Q[k]) * (P[k] - Q[k]);
  }
  p->PA_x[0] = P[0] - A[0]; p->PA_y[0] = P[1] - A[1]; p->PA_z[0] = P[2] - A[2];
  p->PB_x[0] = P[0] - B[0]; p->PB_y[0] = P[1] - B[1]; p->PB_z[0] = P[2] - B[2];
  p->QC_x[0] = Q[0] - C[0]; p->QC_y[0] = Q[1] - C[1]; p->QC_z[0] = Q[2] - C[2];
  p->QD_x[0] = Q[0] - D[0]; p->QD_y[0] = Q[1] - D[1]; p->QD_z[0] = Q[2] - D[2];
  p->AB_x[0] = A[0] - B[0]; p->AB_y[0] = A[1] - B[1]; p->AB_z[0] = A[2] - B[2];
  p->CD_x[0] = C[0] - D[0]; p->CD_y[0] = C[1] - D[1]; p->CD_z[0] = C[2] - D[2];
  p->WP_x[0] = W[0] - P[0]; p->WP_y[0] = W[1] - P[1]; p->WP_z[0] = W[2] - P[2];
  p->WQ_x[0] = W[0] - Q[0]; p->WQ_y[0] = W[1] - Q[1]; p->WQ_z[0] = W[2] - Q[2];
  p->oo2z[0] = 0.5 / gp; p->oo2e[0] = 0.5 / gq; p->oo2ze[0] = 0.5 / gpq;
  p->roz[0] = rho / gp; p->roe[0] = rho / gq;
  pfac = 2 * pow(M_PI, 2.5) * exp(-a1 * a2 * AB2 / gp) * exp(-a3 * a4 * CD2 / gq) / (gp * gq * sqrt(gpq)) * c;
  boys(F, PQ2 * rho, L);
  for (m = 0; m <= L; ++m) SET_F(p, m, pfac * F[m]);
}

int main(void) {
  /* fixed geometry of tests/unit/test-c-api.cc:46-62, exponents extended to a 2-term contraction */
  const double cen[4][3] = {{0.0, 1.0, 2.0}, {1.0, 2.0, 0.0}, {2.0, 0.0, 1.0}, {0.0, 1.0, 2.0}};
  const double al[4][2] = {{1.1, 0.4}, {2.3, 0.7}, {3.4, 0.9}, {4.8, 0.6}};
  const double co[4][2] = {{1.0, 0.5}, {1.0, 0.8}, {1.0, 0.3}, {1.0, 0.6}};
  const int max_am = 2, K = 2;
  Libint_t* ev = (Libint_t*)malloc(sizeof(Libint_t) * K * K * K * K);
  int la, lb, lc, ld;
  libint2_static_init();
  libint2_init_eri(&ev[0], max_am, 0);
  for (la = 0; la <= max_am; ++la)
    for (lb = 0; lb <= la; ++lb)
      for (lc = 0; lc <= max_am; ++lc)
        for (ld = 0; ld <= lc; ++ld) {
          shell_t s[4];
          const int l[4] = {la, lb, lc, ld};
          int i, k, n, ip[4], np = 0, L = la + lb + lc + ld;
          double* out;
          if (la + lb > lc + ld || L == 0) continue; /* canonical classes (build_libint.cc:78-83) */
          if (!libint2_build_eri[la][lb][lc][ld]) {
            printf("class %d %d %d %d missing\n", la, lb, lc, ld);
            return 2;
          }
          for (i = 0; i < 4; ++i) {
            s[i].l = l[i]; s[i].K = K;
            for (k = 0; k < K; ++k) { s[i].alpha[k] = al[i][k]; s[i].coef[k] = co[i][k]; }
            for (k = 0; k < 3; ++k) s[i].O[k] = cen[i][k];
          }
          for (ip[0] = 0; ip[0] < K; ++ip[0])
            for (ip[1] = 0; ip[1] < K; ++ip[1])
              for (ip[2] = 0; ip[2] < K; ++ip[2])
                for (ip[3] = 0; ip[3] < K; ++ip[3]) fill_primitive(&ev[np++], s, ip, L);
          ev[0].contrdepth = np;
          libint2_build_eri[la][lb][lc][ld](&ev[0]);
          out = ev[0].targets[0];
          n = ((la + 1) * (la + 2) / 2) * ((lb + 1) * (lb + 2) / 2) * ((lc + 1) * (lc + 2) / 2) * ((ld + 1) * (ld + 2) / 2);
          printf("class %d %d %d %d n %d\n", la, lb, lc, ld, n);
          for (i = 0; i < n; ++i) printf("%.17g\n", out[i]);
        }
  libint2_cleanup_eri(&ev[0]);
  libint2_static_cleanup();
  free(ev);
  printf("done\n");
  return 0;
}
