/*
 * ORACLE - TEST INFRASTRUCTURE ONLY (see oracle/pandapower_nr.py). Plain-C restatement of the
 * Newton-Raphson power flow that pandapower 2.7.0's default runpp performs for a MAPDN net
 * (PYPOWER newtonpf): polar form, flat start, convergence test before the first
 * solve, <= max_it solves, dense Jacobian solved by LU with partial pivoting (SuperLU's job in the
 * reference). Independent of both the NumPy restatement (sparse dS/dV formulas + spsolve) and the CUDA
 * path (tree elimination): a third implementation of the same iteration.
 *
 * Built by oracle/c_oracle.py (gcc -O2 -shared); only tests/ and bench.py's CPU leg may load it.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* G, B: dense Ybus [n*n] row-major (real / imaginary part). P, Q: specified injections (Sbus) p.u. [n].
 * Returns the number of linear solves; *converged = 1/0; vm, va (rad) [n] out. */
int mapdn_oracle_nr_dense(int n, int slack, const double* G, const double* B, const double* P, const double* Q,
                          double vm_init, double slack_vm, double slack_va, double tol, int max_it,
                          double* vm, double* va, int* converged) {
  const int m = 2 * (n - 1), ld = m + 1;
  int* pq = (int*)malloc(sizeof(int) * (size_t)n);
  int* pos = (int*)malloc(sizeof(int) * (size_t)n);
  double* J = (double*)malloc(sizeof(double) * (size_t)m * (size_t)ld);
  double* Pc = (double*)malloc(sizeof(double) * (size_t)n);
  double* Qc = (double*)malloc(sizeof(double) * (size_t)n);
  int k = 0, it = 0;
  for (int i = 0; i < n; ++i) { pos[i] = -1; if (i != slack) { pq[k] = i; pos[i] = k++; } }
  for (int i = 0; i < n; ++i) { vm[i] = (i == slack) ? slack_vm : vm_init; va[i] = (i == slack) ? slack_va : 0.0; }
  *converged = 0;
  for (;;) {
    double nrm = 0.0;
    for (int i = 0; i < n; ++i) {                     /* S = V conj(Ybus V) */
      double p = 0.0, q = 0.0;
      for (int j = 0; j < n; ++j) {
        const double g = G[(size_t)i * n + j], b = B[(size_t)i * n + j];
        if (g == 0.0 && b == 0.0) continue;
        const double th = va[i] - va[j], c = cos(th), s = sin(th), vv = vm[i] * vm[j];
        p += vv * (g * c + b * s);
        q += vv * (g * s - b * c);
      }
      Pc[i] = p; Qc[i] = q;
      if (i != slack) {
        const double fp = fabs(p - P[i]), fq = fabs(q - Q[i]);
        if (fp > nrm || fp != fp) nrm = fp;
        if (fq > nrm || fq != fq) nrm = fq;
      }
    }
    if (nrm < tol) { *converged = 1; break; }
    if (it >= max_it) break;
    ++it;
    /* J = [[dP/dth, dP/dV],[dQ/dth, dQ/dV]] (unknown order: all angles, then all magnitudes - as newtonpf) */
    memset(J, 0, sizeof(double) * (size_t)m * (size_t)ld);
    const int npq = n - 1;
    for (int a = 0; a < npq; ++a) {
      const int i = pq[a];
      double* rp = J + (size_t)a * ld;
      double* rq = J + (size_t)(npq + a) * ld;
      for (int j = 0; j < n; ++j) {
        if (j == slack) continue;
        const int c = pos[j];
        const double g = G[(size_t)i * n + j], b = B[(size_t)i * n + j];
        if (i == j) {
          rp[c] = -Qc[i] - b * vm[i] * vm[i];
          rp[npq + c] = Pc[i] / vm[i] + g * vm[i];
          rq[c] = Pc[i] - g * vm[i] * vm[i];
          rq[npq + c] = Qc[i] / vm[i] - b * vm[i];
        } else if (g != 0.0 || b != 0.0) {
          const double th = va[i] - va[j], cs = cos(th), sn = sin(th);
          rp[c] = vm[i] * vm[j] * (g * sn - b * cs);
          rp[npq + c] = vm[i] * (g * cs + b * sn);
          rq[c] = -vm[i] * vm[j] * (g * cs + b * sn);
          rq[npq + c] = vm[i] * (g * sn - b * cs);
        }
      }
      rp[m] = -(Pc[i] - P[i]);
      rq[m] = -(Qc[i] - Q[i]);
    }
    /* LU with partial pivoting on [J | rhs] */
    int singular = 0;
    for (int c = 0; c < m && !singular; ++c) {
      int piv = c; double best = fabs(J[(size_t)c * ld + c]);
      for (int r = c + 1; r < m; ++r) { const double v = fabs(J[(size_t)r * ld + c]); if (v > best) { best = v; piv = r; } }
      if (!(best > 0.0)) { singular = 1; break; }
      if (piv != c)
        for (int x = c; x <= m; ++x) { const double t = J[(size_t)c * ld + x]; J[(size_t)c * ld + x] = J[(size_t)piv * ld + x]; J[(size_t)piv * ld + x] = t; }
      const double pinv = 1.0 / J[(size_t)c * ld + c];
      for (int r = c + 1; r < m; ++r) {
        const double l = J[(size_t)r * ld + c] * pinv;
        if (l != 0.0) for (int x = c + 1; x <= m; ++x) J[(size_t)r * ld + x] -= l * J[(size_t)c * ld + x];
      }
    }
    if (singular) break;
    for (int r = m - 1; r >= 0; --r) {
      double acc = J[(size_t)r * ld + m];
      for (int x = r + 1; x < m; ++x) acc -= J[(size_t)r * ld + x] * J[(size_t)x * ld + m];
      J[(size_t)r * ld + m] = acc / J[(size_t)r * ld + r];
    }
    for (int a = 0; a < npq; ++a) { va[pq[a]] += J[(size_t)a * ld + m]; vm[pq[a]] += J[(size_t)(npq + a) * ld + m]; }
  }
  free(pq); free(pos); free(J); free(Pc); free(Qc);
  return it;
}
