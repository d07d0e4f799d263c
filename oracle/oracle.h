/* oracle.h — C entry points of the CPU oracle (TEST INFRASTRUCTURE ONLY, see lentil_oracle.cpp).
 * Same contracts as include/lentil_b200.h, every pointer a HOST pointer. */
#ifndef LENTIL_ORACLE_H
#define LENTIL_ORACLE_H
#include "../include/lentil_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct orc_camera orc_camera;
int orc_camera_create(const lb_camera_params *params, const lb_bokeh_image *bokeh, orc_camera **out);
void orc_camera_destroy(orc_camera *c);
int orc_camera_get_state(const orc_camera *c, lb_camera_state *s);
int orc_camera_set_state(orc_camera *c, double aperture_radius, double sensor_shift);
int orc_camera_create_rays(orc_camera *cam, size_t n, uint64_t ray_id_base, const lb_ray_in *in, const lb_ray_out *out, int nthreads);
int orc_filter_begin(orc_camera *c, const lb_frame_desc *f, int n_aov, const lb_aov_desc *aovs);
int orc_filter_accumulate(orc_camera *cam, const lb_samples *S, int nthreads);
int orc_filter_get_stats(orc_camera *c, lb_filter_stats *out);
int orc_imager_resolve(orc_camera *c, int aov, int x0, int y0, int w, int h, float *rgba_out);
int orc_filter_buffers(orc_camera *c, int aov, float **buffer, float **weight);
int orc_filter_crypto(orc_camera *c, int aov, int slots, float *ids_out, float *weights_out, float *total_out);
void orc_camera_counters(const orc_camera *c, uint64_t out[4]);
#ifdef __cplusplus
}
#endif
#endif
