// stubs.cu — entry points not implemented yet (each returns PITT_ERR_INVALID with a message).
#include "pitt_common.cuh"
using namespace pitt;
extern "C" {
#define NOT_YET(ctx, name) (ctx ? fail(ctx, PITT_ERR_INVALID, name ": not implemented yet") : PITT_ERR_CUDA)
int pitt_find_supports(pitt_ctx* ctx, const pitt_cloud*, const pitt_support_params*, pitt_support_result*) { return NOT_YET(ctx, "pitt_find_supports"); }
int pitt_cluster_service(pitt_ctx* ctx, const pitt_cloud*, const pitt_cluster_params*, pitt_clusters_result*) { return NOT_YET(ctx, "pitt_cluster_service"); }
int pitt_primitive_service(pitt_ctx* ctx, const pitt_cloud*, const pitt_sac_params*, pitt_primitive_result*) { return NOT_YET(ctx, "pitt_primitive_service"); }
int pitt_select_primitive(int64_t, int64_t, int64_t, int64_t, float) { return 0; }
int pitt_segment_frame(pitt_ctx* ctx, const pitt_cloud*, const pitt_frame_params*, pitt_frame_result*) { return NOT_YET(ctx, "pitt_segment_frame"); }
}
