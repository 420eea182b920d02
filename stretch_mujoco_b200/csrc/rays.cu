// Ray casting kernels of libstretchsim: lidar (S2/L1) and pinhole camera (C1-C4).
#include "host.h"
int ss_rays_model_init(ss_model* M) { M->rm.present = 0; return 0; }
int ss_rays_set_fovy(ss_model* M, const double* fovy, size_t bytes) { return ss_fail("ray geometry not built"); }
extern "C" int ss_model_num_rangefinders(const ss_model* M) { return M ? M->nrange : 0; }
extern "C" int ss_batch_lidar(ss_batch* B, float* out_dev, ss_stream s) { return ss_fail("ray geometry not built"); }
extern "C" int ss_batch_rays(ss_batch* B, int nray, const float* o, const float* d, int gm, int be, float* dist, int32_t* geom, ss_stream s) { return ss_fail("ray geometry not built"); }
extern "C" int ss_batch_render(ss_batch* B, int cam, int W, int H, float fovy, uint8_t* rgb, float* depth, float lim, int e0, int ne, ss_stream s) { return ss_fail("ray geometry not built"); }
