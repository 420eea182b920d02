/* CPU oracle: ray casting + camera (rows S2, C3 of SURVEY.md §8(a)) -- TEST INFRASTRUCTURE ONLY. */
#include "ss_oracle.h"
#include "ss_oracle_internal.h"

double om_ray(const om_model* m, const double* xpos, const double* xmat, const double* geom_xpos,
              const double* geom_xmat, const double* pnt, const double* vec, int groupmask, int bodyexclude,
              int* geomid) {
  (void)m; (void)xpos; (void)xmat; (void)geom_xpos; (void)geom_xmat; (void)pnt; (void)vec; (void)groupmask;
  (void)bodyexclude;
  if (geomid) *geomid = -1;
  return -1;
}
int om_batch_rays(const om_model* m, int nenv, const double* xpos, const double* xquat, int nray,
                  const double* origin, const double* dir, int groupmask, int bodyexclude, double* out_dist,
                  int* out_geom, int nthreads) { return -1; }
int om_batch_render(const om_model* m, int nenv, const double* xpos, const double* xquat, int cam_id, int W, int H,
                    double fovy_deg, unsigned char* rgb, float* depth, int nthreads) { return -1; }
