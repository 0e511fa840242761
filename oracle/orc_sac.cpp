// orc_sac.cpp — CPU ORACLE (test infrastructure only; see orc_math.h header for parity status).
// RandomSampleConsensus::computeModel + SACSegmentation::segment of PCL 1.7.x as driven by
// supports_segmentation_srv.cpp:89-111, plane…:52-67, sphere…:58-73, cylinder…:111-126,
// cone…:112-127 of the reference.
#include "orc_sac.h"

#include <algorithm>

#include "oracle.h"
#include "orc_lm.h"

namespace orc {

int countWithinDistance(const Cloud& c, int model, const Limits& L, double thr, const float* mc) {
  if (!isModelValid(model, L, mc)) return 0;
  Scorer s(model, L, thr, mc);
  int n = 0;
  for (int i = 0; i < c.n; ++i) n += s.inlier(c, i) ? 1 : 0;
  return n;
}

void selectWithinDistance(const Cloud& c, int model, const Limits& L, double thr, const float* mc, std::vector<int>& out) {
  out.clear();
  if (!isModelValid(model, L, mc)) return;
  Scorer s(model, L, thr, mc);
  for (int i = 0; i < c.n; ++i)
    if (s.inlier(c, i)) out.push_back(i);
}

// SampleConsensusModelPlane::optimizeModelCoefficients: computeMeanAndCovarianceMatrix + eigen33.
// The nine sums are DEFINED as exact (DD), rounded once to float (orc_math.h).
static void planeRefine(const Cloud& c, const std::vector<int>& inl, const float* mc, float* out) {
  if (inl.size() < 4) {
    for (int i = 0; i < 4; ++i) out[i] = mc[i];
    return;
  }
  DD acc[9];
  for (int idx : inl) {
    double x = c.xyz[4 * idx], y = c.xyz[4 * idx + 1], z = c.xyz[4 * idx + 2];
    acc[0].add(x * x); acc[1].add(x * y); acc[2].add(x * z);
    acc[3].add(y * y); acc[4].add(y * z); acc[5].add(z * z);
    acc[6].add(x); acc[7].add(y); acc[8].add(z);
  }
  float accu[9], cov[9], cen[4];
  for (int i = 0; i < 9; ++i) accu[i] = acc[i].f();
  covFromAccu(accu, (float)inl.size(), cov, cen);
  float ev, evec[3];
  eigen33(cov, ev, evec);
  V4 o = mk(evec[0], evec[1], evec[2], 0.0f);
  out[0] = o[0]; out[1] = o[1]; out[2] = o[2];
  out[3] = -1.0f * dot4(o, mk(cen[0], cen[1], cen[2], cen[3]));
}

int optimizeModelCoefficients(const Cloud& c, int model, const std::vector<int>& inl, const float* mc, float* refined, int* nfev) {
  if (nfev) *nfev = 0;
  if (model == PITT_MODEL_PLANE) {
    planeRefine(c, inl, mc, refined);
    return 0;
  }
  return lmRefine(c, model, inl, mc, refined, nfev);
}

struct Trace {
  std::vector<int> samples;
  std::vector<float> coeffs;
  std::vector<int> counts;
  std::vector<uint8_t> valid;
};

// The RANSAC loop. In ALL_H mode exactly max_iterations stream positions are scored and the
// earliest arg-max wins (what PCL does when the adaptive stop never fires).
static int ransac(const Cloud& c, const pitt_sac_params& p, const Limits& L, float* best_mc, std::vector<int>& best_sel,
                  pitt_sac_info* info, Trace* tr) {
  const int S = sampleSize(p.model);
  const int NC = coeffCount(p.model);
  PclSampler sampler(c.n);
  int iterations = 0, skipped = 0, pos = 0;
  int best = -INT_MAX, best_pos = -1;
  double k = 1.0;
  const double log_probability = log(1.0 - p.probability);
  const double one_over_indices = 1.0 / (double)c.n;
  const int max_skip = p.max_iterations * 10;
  std::vector<int> sel(S);
  float mc[8];
  const bool all_h = (p.stop == PITT_STOP_ALL_H);
  auto next_sample = [&](int* out) -> bool {
    if (p.sampler == PITT_SAMPLER_REPLAY) {
      if (pos >= p.replay_count) return false;
      for (int i = 0; i < S; ++i) out[i] = p.replay_samples[(size_t)pos * S + i];
      return true;
    }
    return sampler.getSamples(c, p.model, out);
  };
  while (all_h ? (pos < p.max_iterations) : ((double)iterations < k && skipped < max_skip)) {
    if (!next_sample(sel.data())) break;
    int this_pos = pos++;
    bool ok = computeModelCoefficients(c, p.model, sel.data(), L, mc);
    if (tr) {
      tr->samples.insert(tr->samples.end(), sel.begin(), sel.end());
      for (int i = 0; i < 8; ++i) tr->coeffs.push_back(i < NC && ok ? mc[i] : 0.0f);
      tr->valid.push_back(ok ? 1 : 0);
    }
    if (!ok) {
      if (tr) tr->counts.push_back(0);
      ++skipped;
      continue;
    }
    int n = countWithinDistance(c, p.model, L, p.distance_threshold, mc);
    if (tr) tr->counts.push_back(n);
    if (n > best) {
      best = n;
      best_pos = this_pos;
      best_sel = sel;
      for (int i = 0; i < NC; ++i) best_mc[i] = mc[i];
      double w = (double)best * one_over_indices;
      double p_no = 1.0 - pow(w, (double)S);
      p_no = std::max(std::numeric_limits<double>::epsilon(), p_no);
      p_no = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no);
      k = log_probability / log(p_no);
    }
    ++iterations;
    if (!all_h && iterations > p.max_iterations) break;
  }
  if (info) {
    info->iterations = iterations;
    info->skipped = skipped;
    info->hypotheses = pos;
    info->best_hypothesis = best_pos;
    info->best_count = best_pos >= 0 ? best : 0;
  }
  return best_pos;
}

}  // namespace orc

using namespace orc;

extern "C" {

int orc_sac_segment(const float* xyz4, const float* nrm4, int n, const pitt_sac_params* p, int32_t* inliers, int cap,
                    int* n_inliers, float* coeffs, int* n_coeffs, pitt_sac_info* info) {
  Cloud c{xyz4, nrm4, n};
  pitt_sac_info local;
  if (!info) info = &local;
  memset(info, 0, sizeof(*info));
  info->best_hypothesis = -1;
  *n_inliers = 0;
  *n_coeffs = 0;
  if ((p->model == PITT_MODEL_CYLINDER || p->model == PITT_MODEL_CONE) && !nrm4) return PITT_ERR_STATE;
  if (n <= 0) return PITT_OK;
  Limits L = limitsFor(*p);
  float mc[8] = {0};
  std::vector<int> sel;
  int best = ransac(c, *p, L, mc, sel, info, nullptr);
  if (best < 0) return PITT_OK;  // "No solution found": both outputs cleared
  const int NC = coeffCount(p->model);
  std::vector<int> inl;
  selectWithinDistance(c, p->model, L, p->distance_threshold, mc, inl);
  info->n_inliers_model = (int)inl.size();
  for (int i = 0; i < NC; ++i) info->model_coeffs[i] = mc[i];
  float out[8] = {0};
  if (p->optimize) {
    int nfev = 0;
    info->lm_info = optimizeModelCoefficients(c, p->model, inl, mc, out, &nfev);
    info->lm_nfev = nfev;
    selectWithinDistance(c, p->model, L, p->distance_threshold, out, inl);
  } else {
    for (int i = 0; i < NC; ++i) out[i] = mc[i];
  }
  for (int i = 0; i < NC; ++i) coeffs[i] = out[i];
  *n_coeffs = NC;
  *n_inliers = (int)inl.size();
  if ((int)inl.size() > cap) return PITT_ERR_CAPACITY;
  for (size_t i = 0; i < inl.size(); ++i) inliers[i] = inl[i];
  return PITT_OK;
}

int orc_sac_score(const float* xyz4, const float* nrm4, int n, const pitt_sac_params* p, const int32_t* samples, int H,
                  int32_t* counts, float* coeffs8, uint8_t* valid) {
  Cloud c{xyz4, nrm4, n};
  Limits L = limitsFor(*p);
  const int S = sampleSize(p->model), NC = coeffCount(p->model);
  for (int h = 0; h < H; ++h) {
    float mc[8] = {0};
    bool ok = computeModelCoefficients(c, p->model, samples + (size_t)h * S, L, mc);
    if (valid) valid[h] = ok ? 1 : 0;
    if (coeffs8)
      for (int i = 0; i < 8; ++i) coeffs8[(size_t)h * 8 + i] = (ok && i < NC) ? mc[i] : 0.0f;
    if (counts) counts[h] = ok ? countWithinDistance(c, p->model, L, p->distance_threshold, mc) : 0;
  }
  return PITT_OK;
}

int orc_sac_select(const float* xyz4, const float* nrm4, int n, const pitt_sac_params* p, const float* coeffs,
                   int32_t* inliers, int cap, int* n_inliers) {
  Cloud c{xyz4, nrm4, n};
  Limits L = limitsFor(*p);
  std::vector<int> inl;
  selectWithinDistance(c, p->model, L, p->distance_threshold, coeffs, inl);
  *n_inliers = (int)inl.size();
  if ((int)inl.size() > cap) return PITT_ERR_CAPACITY;
  for (size_t i = 0; i < inl.size(); ++i) inliers[i] = inl[i];
  return PITT_OK;
}

int orc_sac_refine(const float* xyz4, const float* nrm4, int n, const pitt_sac_params* p, const float* coeffs,
                   const int32_t* inliers, int n_inliers, float* refined, pitt_sac_info* info) {
  Cloud c{xyz4, nrm4, n};
  std::vector<int> inl(inliers, inliers + n_inliers);
  int nfev = 0;
  int r = optimizeModelCoefficients(c, p->model, inl, coeffs, refined, &nfev);
  if (info) {
    info->lm_info = r;
    info->lm_nfev = nfev;
  }
  return PITT_OK;
}

int orc_pcl_sample_stream(const float* xyz4, int n, int model, int count, int32_t* out) {
  Cloud c{xyz4, nullptr, n};
  PclSampler s(n);
  const int S = sampleSize(model);
  for (int h = 0; h < count; ++h)
    if (!s.getSamples(c, model, out + (size_t)h * S)) return PITT_ERR_INVALID;
  return PITT_OK;
}

// raw mt19937 words (known-answer check: the 10000th output of mt19937(5489) is 4123659995)
uint32_t orc_mt19937_nth(uint32_t seed, int nth) {
  MT19937 mt(seed);
  uint32_t v = 0;
  for (int i = 0; i < nth; ++i) v = mt.next();
  return v;
}

}  // extern "C"
