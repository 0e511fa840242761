// sac.cuh — host-side interface of sac.cu / lm.cu used by the C ABI and the service layer.
#pragma once
#include "pitt_common.cuh"
#include "sac_device.cuh"

namespace pitt {

constexpr int REF_TPB = 256;     // plane refinement: fixed reduction shape (the partial sums of a problem take REF_BLOCKS x 10 doubles)
constexpr int REF_BLOCKS = 296;

struct SacDeviceResult {
  pitt_sac_info info;
  float coeffs[8];
  int n_coeffs;
  int n_inliers;
  const int* d_inliers;  // device, ascending, valid until the next outermost API call (arena)
};

// pcl::SampleConsensusModel::getSamples replay (host). h_xyz4 may be null: draws are then
// speculative for the plane model (no isSampleGood redraw).
class PclSampleStream {
 public:
  PclSampleStream(int n, int model, const float* h_xyz4);
  ~PclSampleStream();
  bool next(int* out);

 private:
  struct Impl;
  Impl* impl_;
  int model_;
  const float* h_xyz_;
};

Limits limits_for(const pitt_sac_params& p);
ScoreParams score_params_for(const pitt_sac_params& p, const Limits& L);

int sac_estimate(pitt_ctx* ctx, const pitt_cloud* c, int model, const int* d_samples, int H, const Limits& L, const ScoreParams& sp,
                 HypRec* d_recs, float* d_coeffs8, uint8_t* d_flags);
int sac_score(pitt_ctx* ctx, const pitt_cloud* c, int model, const HypRec* d_recs, int H, const ScoreParams& sp,
              int* d_counts);
int sac_winner(pitt_ctx* ctx, const int* d_counts, const uint8_t* d_flags, int H, const float* d_coeffs8, int* d_best,
               float* d_best_coeffs);
int sac_select(pitt_ctx* ctx, const pitt_cloud* c, int model, const float* d_coeffs, const Limits& L,
               const ScoreParams& sp, int* d_out, int* d_total);
int plane_refine(pitt_ctx* ctx, const pitt_cloud* c, const float* d_model, const int* d_idx, const int* d_n_idx,
                 const Limits& L, const ScoreParams& sp, float* d_refined, int* d_n_model_inliers);
int lm_refine(pitt_ctx* ctx, const pitt_cloud* c, int model, const float* d_model, const int* d_idx, const int* d_n_idx,
              int n_idx_host, float* d_refined, int* d_lm_info);
int sac_philox_samples(pitt_ctx* ctx, int* d_samples, int H, int S, int n, uint32_t stream_id);
int sac_philox_raw(pitt_ctx* ctx, const uint32_t* h_ctr_key6, uint32_t* h_out4);
int sac_finish_from_winner(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params& p, const int* d_samples_all, int H_all,
                           const int* d_best, SacDeviceResult* out);
int sac_segment_impl(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params& p, SacDeviceResult* out);
// seg.segment() enqueued on ctx->stream without any host round trip (sac.cu)
struct SacAsync {
  int* d_ints;   // 16 ints, see sac_segment_async
  float* d_flt;  // model[8], refined[8]
  int* d_inl;    // final inliers (ascending), capacity n
  int H;         // hypotheses scored
};
int sac_segment_async(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params& p, int* h_stage, SacAsync* out, bool* issued);
// The same for a batch of problems that share the model and its parameters (the clusters of a support): every launch serves all
// of them (problem = a grid dimension, pointers and sizes from d_desc). h_desc: the host copy of the descriptors (n, H filled in).
// Every problem must have at most 4096 points and the PCL adaptive stop.
int sac_fit_batch_async(pitt_ctx* ctx, const pitt_sac_params& p, const FitDesc* h_desc, const FitDesc* d_desc, int nprob, bool speculative_plane);
int lm_refine_batch(pitt_ctx* ctx, int model, const FitDesc* d_desc, int nprob, int m_cap);
int fp32_peak(pitt_ctx* ctx, int kind, double* tflops);

extern int g_force_generic_plane;
extern int g_plane_mode;
bool plane_job_takes_tensor_path(int n, int H);
extern int g_score_mode;
extern int g_select_no_fuse;
extern int g_frame_legacy;
extern int g_frame_no_batch;
extern int g_lm_cluster_min;
extern unsigned long long g_plane_filter_stats[2];
extern int g_plane_filter_collect_stats;
extern unsigned long long g_plane_tc_stats[2 + 160 + 16];
extern int g_plane_tc_collect_stats;
extern int g_plane_tc_dump;
extern int g_plane_tc_variant;
extern int g_plane_tc_nwq;
extern float g_plane_tc_acc_ulps;
extern std::vector<float> g_plane_tc_dump_host;

}  // namespace pitt
