// orc_lm.h — CPU ORACLE (test infrastructure only).
// Eigen 3.2 unsupported/NonLinearOptimization LevenbergMarquardt<NumericalDiff<Functor>, float>
// ::minimize() as used by SampleConsensusModel{Sphere,Cylinder,Cone}::optimizeModelCoefficients
// (PCL 1.7.x), restated without Eigen.
//
// PINNED CHOICES (parity unpinned, see orc_math.h):
//  * every reduction over the m residual rows (column norms, Householder dot products, Q^T f,
//    ||f||) is DEFINED as the exact sum rounded once to float (DD accumulator) — Eigen's packet
//    order is build dependent and a parallel reduction cannot follow a sequential one;
//  * every n-sized (n <= 7) loop runs left to right in float;
//  * blueNorm()/stableNorm() = sqrtf(sum of squares) (the overflow guards never trigger here);
//  * tanf in the cone functor = tan in double rounded to float (orc_math.h).
#pragma once
#include <algorithm>
#include <vector>

#include "orc_sac.h"

namespace orc {

struct LMFunctor {
  const Cloud* c;
  const std::vector<int>* idx;
  int model;
  int n() const { return model == PITT_MODEL_SPHERE ? 4 : 7; }
  int m() const { return (int)idx->size(); }
  void operator()(const float* x, float* fvec) const {
    const int M = m();
    if (model == PITT_MODEL_SPHERE) {
      for (int i = 0; i < M; ++i) {
        int p = (*idx)[i];
        V4 cen = mk(c->xyz[4 * p] - x[0], c->xyz[4 * p + 1] - x[1], c->xyz[4 * p + 2] - x[2], 0.0f);
        fvec[i] = sqrtf(dot4(cen, cen)) - x[3];
      }
    } else if (model == PITT_MODEL_CYLINDER) {
      V4 lp = mk(x[0], x[1], x[2], 0), ld = mk(x[3], x[4], x[5], 0);
      for (int i = 0; i < M; ++i) {
        V4 pt = c->p((*idx)[i]);
        fvec[i] = (float)(sqrPointToLineDistance(pt, lp, ld) - (double)(x[6] * x[6]));
      }
    } else {
      V4 apex = mk(x[0], x[1], x[2], 0), ad = mk(x[3], x[4], x[5], 0);
      float apexdotdir = dot4(apex, ad);
      float dirdotdir = 1.0f / dot4(ad, ad);
      float tan_a = tanf_d(x[6]);
      for (int i = 0; i < M; ++i) {
        V4 pt = c->p((*idx)[i]);
        float k = (dot4(pt, ad) - apexdotdir) * dirdotdir;
        V4 proj = apex + k * ad;
        V4 height = apex - proj;
        float r = tan_a * norm4(height);
        fvec[i] = (float)(sqrPointToLineDistance(pt, apex, ad) - (double)(r * r));
      }
    }
  }
};

// exact sum of products over rows [r0, m), rounded once to float
inline float dotM(const float* a, const float* b, int r0, int m) {
  DD s;
  for (int i = r0; i < m; ++i) s.add((double)a[i] * (double)b[i]);
  return s.f();
}
inline float normM(const float* a, int m) { return sqrtf(dotM(a, a, 0, m)); }
inline float normN(const float* a, int n) {
  float s = 0.0f;
  for (int i = 0; i < n; ++i) s += a[i] * a[i];
  return sqrtf(s);
}

struct LMState {
  int n, m;
  std::vector<float> fjac;  // n columns of m
  float R(int i, int j) const { return fjac[(size_t)j * m + i]; }
  float* col(int j) { return fjac.data() + (size_t)j * m; }
  float hcoef[8];
  int perm[8];  // colsPermutation().indices()
  int nonzero_pivots;
  float maxpivot;
};

// Eigen::ColPivHouseholderQR<MatrixXf>::compute (Eigen 3.2)
inline void colPivQR(LMState& S) {
  const int n = S.n, m = S.m;
  float colSq[8];
  int transp[8];
  for (int k = 0; k < n; ++k) colSq[k] = dotM(S.col(k), S.col(k), 0, m);
  float mx = colSq[0];
  for (int k = 1; k < n; ++k) mx = std::max(mx, colSq[k]);
  const float eps = std::numeric_limits<float>::epsilon();
  float threshold_helper = mx * (eps * eps) / (float)m;
  S.nonzero_pivots = n;
  S.maxpivot = 0.0f;
  for (int k = 0; k < n; ++k) {
    int big = k;
    for (int j = k + 1; j < n; ++j)
      if (colSq[j] > colSq[big]) big = j;
    float bigSq = dotM(S.col(big), S.col(big), k, m);
    colSq[big] = bigSq;
    if (bigSq < threshold_helper * (float)(m - k)) {
      S.nonzero_pivots = k;
      for (int j = k; j < n; ++j) {
        S.hcoef[j] = 0.0f;
        transp[j] = j;
        float* cj = S.col(j);
        for (int i = std::max(k, j) + 1; i < m; ++i)
          if (i > j) cj[i] = 0.0f;  // strictly lower part of the bottom-right corner
      }
      break;
    }
    transp[k] = big;
    if (k != big) {
      float* a = S.col(k);
      float* b = S.col(big);
      for (int i = 0; i < m; ++i) std::swap(a[i], b[i]);
      std::swap(colSq[k], colSq[big]);
    }
    // makeHouseholderInPlace on col k, rows k..m-1
    float* ck = S.col(k);
    float tailSq = (m - k == 1) ? 0.0f : dotM(ck, ck, k + 1, m);
    float c0 = ck[k];
    float tau, beta;
    if (tailSq == 0.0f) {
      tau = 0.0f;
      beta = c0;
      for (int i = k + 1; i < m; ++i) ck[i] = 0.0f;
    } else {
      beta = sqrtf(c0 * c0 + tailSq);
      if (c0 >= 0.0f) beta = -beta;
      float den = c0 - beta;
      for (int i = k + 1; i < m; ++i) ck[i] = ck[i] / den;
      tau = (beta - c0) / beta;
    }
    S.hcoef[k] = tau;
    ck[k] = beta;
    if (fabsf(beta) > S.maxpivot) S.maxpivot = fabsf(beta);
    // applyHouseholderOnTheLeft to columns k+1..n-1, rows k..m-1
    for (int j = k + 1; j < n; ++j) {
      float* cj = S.col(j);
      if (m - k == 1) {
        cj[k] *= (1.0f - tau);
      } else {
        float tmp = dotM(ck, cj, k + 1, m);
        tmp += cj[k];
        cj[k] -= tau * tmp;
        for (int i = k + 1; i < m; ++i) cj[i] -= tmp * (tau * ck[i]);
      }
    }
    for (int j = k + 1; j < n; ++j) colSq[j] -= S.col(j)[k] * S.col(j)[k];
  }
  for (int j = 0; j < n; ++j) S.perm[j] = j;
  for (int k = 0; k < S.nonzero_pivots; ++k) std::swap(S.perm[k], S.perm[transp[k]]);
}

// w <- Q^T w (householderQ().adjoint() applied on the left)
inline void applyQT(LMState& S, float* w) {
  const int n = S.n, m = S.m;
  for (int k = 0; k < n; ++k) {
    float tau = S.hcoef[k];
    const float* ck = S.col(k);
    if (m - k == 1) {
      w[k] *= (1.0f - tau);
    } else {
      float tmp = dotM(ck, w, k + 1, m);
      tmp += w[k];
      w[k] -= tau * tmp;
      for (int i = k + 1; i < m; ++i) w[i] -= tmp * (tau * ck[i]);
    }
  }
}

inline void makeGivens(float p, float q, float& c, float& s) {
  if (q == 0.0f) {
    c = p < 0.0f ? -1.0f : 1.0f;
    s = 0.0f;
  } else if (p == 0.0f) {
    c = 0.0f;
    s = q < 0.0f ? 1.0f : -1.0f;
  } else if (fabsf(p) > fabsf(q)) {
    float t = q / p;
    float u = sqrtf(1.0f + t * t);
    if (p < 0.0f) u = -u;
    c = 1.0f / u;
    s = -t * c;
  } else {
    float t = p / q;
    float u = sqrtf(1.0f + t * t);
    if (q < 0.0f) u = -u;
    s = -1.0f / u;
    c = -t * s;
  }
}

// Eigen internal::qrsolv on the n x n matrix s (row i, col j at s[i*8+j])
inline void qrsolv(float s[64], int n, const int* ipvt, const float* diag, const float* qtb, float* x, float* sdiag) {
  float wa[8];
  for (int j = 0; j < n; ++j) { x[j] = s[j * 8 + j]; wa[j] = qtb[j]; }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) s[i * 8 + j] = s[j * 8 + i];
  for (int j = 0; j < n; ++j) {
    int l = ipvt[j];
    if (diag[l] == 0.0f) break;
    for (int k = j; k < n; ++k) sdiag[k] = 0.0f;
    sdiag[j] = diag[l];
    float qtbpj = 0.0f;
    for (int k = j; k < n; ++k) {
      float gc, gs;
      makeGivens(-s[k * 8 + k], sdiag[k], gc, gs);
      s[k * 8 + k] = gc * s[k * 8 + k] + gs * sdiag[k];
      float temp = gc * wa[k] + gs * qtbpj;
      qtbpj = -gs * wa[k] + gc * qtbpj;
      wa[k] = temp;
      for (int i = k + 1; i < n; ++i) {
        temp = gc * s[i * 8 + k] + gs * sdiag[i];
        sdiag[i] = -gs * s[i * 8 + k] + gc * sdiag[i];
        s[i * 8 + k] = temp;
      }
    }
  }
  int nsing;
  for (nsing = 0; nsing < n && sdiag[nsing] != 0.0f; nsing++) {}
  for (int j = nsing; j < n; ++j) wa[j] = 0.0f;
  // s.topLeftCorner(nsing,nsing).transpose().triangularView<Upper>().solveInPlace(wa)
  for (int i = nsing - 1; i >= 0; --i) {
    float acc = 0.0f;
    for (int j = i + 1; j < nsing; ++j) acc += s[j * 8 + i] * wa[j];
    wa[i] = (wa[i] - acc) / s[i * 8 + i];
  }
  for (int j = 0; j < n; ++j) { sdiag[j] = s[j * 8 + j]; s[j * 8 + j] = x[j]; }
  for (int j = 0; j < n; ++j) x[ipvt[j]] = wa[j];
}

// Eigen internal::lmpar2. r = top n x n of the QR factor (row i, col j at r[i*8+j]).
inline void lmpar2(const float r[64], int n, const int* perm, int rank, const float* diag, const float* qtb, float delta,
                   float& par, float* x) {
  const float dwarf = std::numeric_limits<float>::min();
  float wa1[8], wa2[8];
  for (int j = 0; j < n; ++j) wa1[j] = qtb[j];
  for (int j = rank; j < n; ++j) wa1[j] = 0.0f;
  // column-major upper triangular solve
  for (int i = rank - 1; i >= 0; --i) {
    wa1[i] /= r[i * 8 + i];
    for (int q = 0; q < i; ++q) wa1[q] -= wa1[i] * r[q * 8 + i];
  }
  for (int j = 0; j < n; ++j) x[perm[j]] = wa1[j];
  int iter = 0;
  for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
  float dxnorm = normN(wa2, n);
  float fp = dxnorm - delta;
  if (fp <= 0.1f * delta) {
    par = 0.0f;
    return;
  }
  float parl = 0.0f;
  if (rank == n) {
    for (int j = 0; j < n; ++j) wa1[j] = diag[perm[j]] * wa2[perm[j]] / dxnorm;
    // R^T lower, row oriented forward substitution
    for (int i = 0; i < n; ++i) {
      float acc = 0.0f;
      for (int j = 0; j < i; ++j) acc += r[j * 8 + i] * wa1[j];
      wa1[i] = (wa1[i] - acc) / r[i * 8 + i];
    }
    float temp = normN(wa1, n);
    parl = fp / delta / temp / temp;
  }
  for (int j = 0; j < n; ++j) {
    float acc = 0.0f;
    for (int i = 0; i <= j; ++i) acc += r[i * 8 + j] * qtb[i];
    wa1[j] = acc / diag[perm[j]];
  }
  float gnorm = normN(wa1, n);
  float paru = gnorm / delta;
  if (paru == 0.0f) paru = dwarf / std::min(delta, 0.1f);
  par = std::max(par, parl);
  par = std::min(par, paru);
  if (par == 0.0f) par = gnorm / dxnorm;
  float s[64];
  for (int i = 0; i < 64; ++i) s[i] = r[i];
  float sdiag[8];
  while (true) {
    ++iter;
    if (par == 0.0f) par = std::max(dwarf, 0.001f * paru);
    float sp = sqrtf(par);
    for (int j = 0; j < n; ++j) wa1[j] = sp * diag[j];
    qrsolv(s, n, perm, wa1, qtb, x, sdiag);
    for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = normN(wa2, n);
    float temp = fp;
    fp = dxnorm - delta;
    if (fabsf(fp) <= 0.1f * delta || (parl == 0.0f && fp <= temp && temp < 0.0f) || iter == 10) break;
    for (int j = 0; j < n; ++j) wa1[j] = diag[perm[j]] * (wa2[perm[j]] / dxnorm);
    for (int j = 0; j < n; ++j) {
      wa1[j] /= sdiag[j];
      temp = wa1[j];
      for (int i = j + 1; i < n; ++i) wa1[i] -= s[i * 8 + j] * temp;
    }
    temp = normN(wa1, n);
    float parc = fp / delta / temp / temp;
    if (fp > 0.0f) parl = std::max(parl, par);
    if (fp < 0.0f) paru = std::min(paru, par);
    par = std::max(parl, par + parc);
  }
  if (iter == 0) par = 0.0f;
}

// LevenbergMarquardt::minimize. Returns the Eigen status code.
inline int lmMinimize(const LMFunctor& F, float* x, int* nfev_out) {
  const int n = F.n(), m = F.m();
  const float eps = std::numeric_limits<float>::epsilon();
  const float ftol = sqrtf(eps), xtol = sqrtf(eps), gtol = 0.0f, factor = 100.0f;
  const int maxfev = 400;
  int nfev = 0;
  if (nfev_out) *nfev_out = 0;
  if (n <= 0 || m < n) return 0;  // ImproperInputParameters
  LMState S;
  S.n = n;
  S.m = m;
  S.fjac.assign((size_t)n * m, 0.0f);
  std::vector<float> fvec(m), wa4(m), val2(m);
  float diag[8], qtf[8], wa1[8], wa2[8], wa3[8];
  nfev = 1;
  F(x, fvec.data());
  float fnorm = normM(fvec.data(), m);
  float par = 0.0f, delta = 0.0f, xnorm = 0.0f;
  int iter = 1;
  int status = -1;
  while (status == -1) {
    // --- NumericalDiff forward: re-evaluates f(x), then n perturbed evaluations
    {
      const float h_eps = sqrtf(eps);
      std::vector<float>& val1 = wa4;  // scratch
      F(x, val1.data());
      nfev++;
      float xs[8];
      for (int j = 0; j < n; ++j) xs[j] = x[j];
      for (int j = 0; j < n; ++j) {
        float h = h_eps * fabsf(xs[j]);
        if (h == 0.0f) h = h_eps;
        xs[j] += h;
        F(xs, val2.data());
        nfev++;
        xs[j] = x[j];
        float* cj = S.col(j);
        for (int i = 0; i < m; ++i) cj[i] = (val2[i] - val1[i]) / h;
      }
    }
    for (int j = 0; j < n; ++j) wa2[j] = normM(S.col(j), m);
    colPivQR(S);
    if (iter == 1) {
      for (int j = 0; j < n; ++j) diag[j] = (wa2[j] == 0.0f) ? 1.0f : wa2[j];
      float t[8];
      for (int j = 0; j < n; ++j) t[j] = diag[j] * x[j];
      xnorm = normN(t, n);
      delta = factor * xnorm;
      if (delta == 0.0f) delta = factor;
    }
    for (int i = 0; i < m; ++i) wa4[i] = fvec[i];
    applyQT(S, wa4.data());
    for (int j = 0; j < n; ++j) qtf[j] = wa4[j];
    float r[64];
    for (int i = 0; i < 64; ++i) r[i] = 0.0f;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) r[i * 8 + j] = S.R(i, j);
    float gnorm = 0.0f;
    if (fnorm != 0.0f)
      for (int j = 0; j < n; ++j)
        if (wa2[S.perm[j]] != 0.0f) {
          float acc = 0.0f;
          for (int i = 0; i <= j; ++i) acc += r[i * 8 + j] * (qtf[i] / fnorm);
          gnorm = std::max(gnorm, fabsf(acc / wa2[S.perm[j]]));
        }
    if (gnorm <= gtol) { status = 4; break; }
    for (int j = 0; j < n; ++j) diag[j] = std::max(diag[j], wa2[j]);
    // qr.rank()
    int rank = 0;
    {
      float thr = fabsf(S.maxpivot) * (eps * (float)n);
      for (int i = 0; i < S.nonzero_pivots; ++i) rank += (fabsf(r[i * 8 + i]) > thr) ? 1 : 0;
    }
    float ratio = 0.0f;
    do {
      lmpar2(r, n, S.perm, rank, diag, qtf, delta, par, wa1);
      for (int j = 0; j < n; ++j) { wa1[j] = -wa1[j]; wa2[j] = x[j] + wa1[j]; }
      float t[8];
      for (int j = 0; j < n; ++j) t[j] = diag[j] * wa1[j];
      float pnorm = normN(t, n);
      if (iter == 1) delta = std::min(delta, pnorm);
      F(wa2, wa4.data());
      ++nfev;
      float fnorm1 = normM(wa4.data(), m);
      float actred = -1.0f;
      if (0.1f * fnorm1 < fnorm) { float q = fnorm1 / fnorm; actred = 1.0f - q * q; }
      // wa3 = R * (P^-1 wa1)
      for (int i = 0; i < n; ++i) {
        float acc = 0.0f;
        for (int j = i; j < n; ++j) acc += r[i * 8 + j] * wa1[S.perm[j]];
        wa3[i] = acc;
      }
      float q1 = normN(wa3, n) / fnorm;
      float temp1 = q1 * q1;
      float q2 = sqrtf(par) * pnorm / fnorm;
      float temp2 = q2 * q2;
      float prered = temp1 + temp2 / 0.5f;
      float dirder = -(temp1 + temp2);
      ratio = 0.0f;
      if (prered != 0.0f) ratio = actred / prered;
      if (ratio <= 0.25f) {
        float temp = 0.0f;
        if (actred >= 0.0f) temp = 0.5f;
        if (actred < 0.0f) temp = 0.5f * dirder / (dirder + 0.5f * actred);
        if (0.1f * fnorm1 >= fnorm || temp < 0.1f) temp = 0.1f;
        delta = temp * std::min(delta, pnorm / 0.1f);
        par /= temp;
      } else if (!(par != 0.0f && ratio < 0.75f)) {
        delta = pnorm / 0.5f;
        par = 0.5f * par;
      }
      if (ratio >= 1e-4f) {
        for (int j = 0; j < n; ++j) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
        fvec.swap(wa4);
        xnorm = normN(wa2, n);
        fnorm = fnorm1;
        ++iter;
      }
      bool small_red = fabsf(actred) <= ftol && prered <= ftol && 0.5f * ratio <= 1.0f;
      if (small_red && delta <= xtol * xnorm) { status = 3; break; }
      if (small_red) { status = 1; break; }
      if (delta <= xtol * xnorm) { status = 2; break; }
      if (nfev >= maxfev) { status = 5; break; }
      if (fabsf(actred) <= eps && prered <= eps && 0.5f * ratio <= 1.0f) { status = 6; break; }
      if (delta <= eps * xnorm) { status = 7; break; }
      if (gnorm <= eps) { status = 8; break; }
    } while (ratio < 1e-4f);
  }
  if (nfev_out) *nfev_out = nfev;
  return status;
}

// optimizeModelCoefficients of sphere / cylinder / cone
inline int lmRefine(const Cloud& c, int model, const std::vector<int>& inl, const float* mc, float* out, int* nfev) {
  const int NC = coeffCount(model);
  for (int i = 0; i < NC; ++i) out[i] = mc[i];
  if (nfev) *nfev = 0;
  if (model == PITT_MODEL_SPHERE) {
    if (inl.size() <= 4) return 0;
  } else if (inl.empty()) {
    return 0;
  }
  LMFunctor F{&c, &inl, model};
  int info = lmMinimize(F, out, nfev);
  if (model != PITT_MODEL_SPHERE) {
    // Eigen::Vector3f line_dir(...); line_dir.normalize();
    float nn = sqrtf(out[3] * out[3] + out[4] * out[4] + out[5] * out[5]);
    out[3] /= nn; out[4] /= nn; out[5] /= nn;
  }
  return info;
}

}  // namespace orc
