// orc_params.cpp — CPU ORACLE (test infrastructure only): the reference's launch-parameter defaults.
// plane_segmentation_srv.cpp:19-24, sphere…:19-26, cylinder…:23-30, cone…:24-31,
// supports_segmentation_srv.cpp:35-37 (paths under /root/reference/src/segmentation_services).
#include <cfloat>
#include <cmath>
#include <cstring>

#include "oracle.h"

extern "C" {

void orc_default_sac_params(int model, pitt_sac_params* p) {
  memset(p, 0, sizeof(*p));
  p->model = model;
  p->max_iterations = 1000;
  p->probability = 0.99;
  p->radius_min = -DBL_MAX;
  p->radius_max = DBL_MAX;
  p->min_angle = -DBL_MAX;
  p->max_angle = DBL_MAX;
  p->optimize = 1;
  p->sampler = PITT_SAMPLER_PCL_MT19937;
  p->stop = PITT_STOP_PCL_ADAPTIVE;
  switch (model) {
    case PITT_MODEL_PLANE:
      p->normal_distance_weight = 0.001; p->distance_threshold = 0.007; p->eps_angle = 0.0;
      p->min_angle = 0.0 / 180.0 * M_PI; p->max_angle = 10.0 / 180.0 * M_PI;
      break;
    case PITT_MODEL_SPHERE:
      p->normal_distance_weight = 0.001; p->distance_threshold = 0.007; p->eps_angle = 0.0;
      p->radius_min = 0.005; p->radius_max = 0.500;
      p->min_angle = 100.0 / 180.0 * M_PI; p->max_angle = 180.0 / 180.0 * M_PI;
      break;
    case PITT_MODEL_CYLINDER:
      p->normal_distance_weight = 0.001; p->distance_threshold = 0.008; p->eps_angle = 0.0001;
      p->radius_min = 0.005; p->radius_max = 0.500;
      p->min_angle = 50.0 / 180.0 * M_PI; p->max_angle = 180.0 / 180.0 * M_PI;
      break;
    default:
      p->normal_distance_weight = 0.0006; p->distance_threshold = 0.0055; p->eps_angle = 0.4;
      p->radius_min = 0.001; p->radius_max = 0.500;
      p->min_angle = 10.0 / 180.0 * M_PI; p->max_angle = 170.0 / 180.0 * M_PI;
      break;
  }
}

void orc_default_support_sac_params(pitt_sac_params* p) {
  orc_default_sac_params(PITT_MODEL_PLANE, p);
  p->distance_threshold = (double)0.02f;    // float global widened by setDistanceThreshold(double)
  p->normal_distance_weight = (double)0.9f;
  p->max_iterations = 10;
  p->min_angle = -DBL_MAX;                  // the supports service never calls setMinMaxOpeningAngle
  p->max_angle = DBL_MAX;
}

}  // extern "C"
