/* ORACLE -- TEST INFRASTRUCTURE ONLY (see pgo_orb.cc header).  C API of liboracle.so, loaded by tests via ctypes. */
#ifndef PGO_H_
#define PGO_H_
#include <stddef.h>
#include <stdint.h>
#include "../include/pgb200.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct pgo_orb pgo_orb;
pgo_orb* pgo_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh);
void pgo_orb_destroy(pgo_orb*);
int pgo_orb_tables(const pgo_orb*, float* scale, float* invScale, float* sigma2, float* invSigma2, int32_t* nPer,
                   int32_t* umax16);
int pgo_orb_level_size(const pgo_orb*, int w, int h, int level, int* lw, int* lh);
int pgo_orb_extract(pgo_orb*, const uint8_t* gray, int w, int h, size_t pitch, pgb_keypoint* kps, uint8_t* desc,
                    int cap);
void pgo_orb_stage_times(pgo_orb*, double* t6, int reset);
int pgo_orb_get_level(const pgo_orb*, int level, uint8_t* out, int* w, int* h);
int pgo_orb_get_candidates(const pgo_orb*, int level, int32_t* xyr, int cap);
int pgo_orb_get_level_keypoints(const pgo_orb*, int level, pgb_keypoint* out, int cap);
void pgo_resize_linear(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
int pgo_fast(const uint8_t* img, int w, int h, int th, int nms, int32_t* xys, int cap);
void pgo_fast_score_map(const uint8_t* img, int w, int h, int min_th, uint8_t* out);
void pgo_gaussian_blur7(const uint8_t* src, int w, int h, uint8_t* dst);
void pgo_to_gray(const uint8_t* src, int w, int h, int channels, int rgb_order, int vflip, int hflip, int formula,
                 uint8_t* dst);
float pgo_fast_atan2(float y, float x);
void pgo_fast_atan2_many(const float* y, const float* x, float* out, int n);
float pgo_ic_angle(const uint8_t* img, int w, int h, int cx, int cy);
void pgo_orb_descriptor(const uint8_t* blurred, int w, int h, int cx, int cy, float angle_deg, uint8_t* desc32);
int pgo_distribute_octree(const int32_t* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int32_t* keep,
                          int cap);
int pgo_descriptor_distance(const uint8_t* a, const uint8_t* b);
int pgo_search_by_projection(const pgb_keypoint* cur_kps, const uint8_t* cur_desc, int n_cur, const float* q_uv,
                             const int32_t* q_octave, const float* q_angle, const uint8_t* q_desc,
                             const uint8_t* q_valid, int n_q, float minX, float maxX, float minY, float maxY, float th,
                             const float* scale_factors, int nlevels, int check_ori, int32_t* match_of_cur,
                             int32_t* best_dist_of_q);
int pgo_search_for_initialization(const pgb_keypoint* k1, const uint8_t* d1, int n1, const pgb_keypoint* k2,
                                  const uint8_t* d2, int n2, float* prev_matched, int windowSize, float minX, float maxX,
                                  float minY, float maxY, float nnratio, int check_ori, int32_t* vnMatches12);
int pgo_search_map_points(const pgb_keypoint* kps, const uint8_t* desc, int n, const uint8_t* has_map_point,
                          const float* proj_xy, const int32_t* track_level, const float* view_cos, const uint8_t* mp_desc,
                          const uint8_t* in_view, const uint8_t* mp_observed, int n_mp, float minX, float maxX, float minY,
                          float maxY, float th, const float* scale_factors, float nnratio, int32_t* match_of_feature);
int pgo_search_by_bow(const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_has_map_point,
                      const uint32_t* kf_node_id, const int32_t* kf_feat_start, const uint32_t* kf_feat_idx, int kf_nodes,
                      const uint8_t* f_desc, const float* f_angle, int f_n, const uint32_t* f_node_id,
                      const int32_t* f_feat_start, const uint32_t* f_feat_idx, int f_nodes, float nnratio, int check_ori,
                      int32_t* match_of_feature);
int pgo_distinctive_descriptor(const uint8_t* desc, int N);
void pgo_pose_se3_oplus(const double* update6, const double* pose7, double* out7);
void pgo_pose_edge(const double* pose7, const double* Xw, const double* obs, double fx, double fy, double cx, double cy, double* err2,
                   double* J12);
void pgo_pose_huber(double delta, double e, double* rho3);
void* pgo_pose_problem_create(const float* Tcw, const float* kp_xy, const int32_t* kp_octave, const float* mp_xyz,
                              const uint8_t* has_map_point, int n, const float* inv_level_sigma2, float fx, float fy, float cx,
                              float cy, int robust);
void pgo_pose_problem_destroy(void* h);
void pgo_pose_problem_reset(void* h, const float* Tcw);
void pgo_pose_problem_get_estimate(void* h, double* pose7);
void pgo_pose_problem_set_estimate(void* h, const double* pose7);
void pgo_pose_problem_compute_active_errors(void* h);
double pgo_pose_problem_active_robust_chi2(void* h);
void pgo_pose_problem_build_system(void* h, double* H36, double* b6);
int pgo_pose_ldlt6_solve(const double* H36, const double* b6, double* x6);
void pgo_pose_problem_oplus(void* h, const double* x6);
void pgo_pose_problem_optimize(void* h, int iterations);
void* pgo_pose_problem_create_raw(int n, const double* obs2, const double* Xw3, const double* info, double fx, double fy, double cx,
                                  double cy, double delta);
void pgo_pose_problem_edge_set_level(void* h, int k, int level);
void pgo_pose_problem_edge_set_robust(void* h, int k, int robust);
void pgo_pose_problem_edge_compute_error(void* h, int k);
double pgo_pose_problem_edge_chi2(void* h, int k);
int pgo_pose_problem_num_active(void* h);
void pgo_pose_T_to_pose7(const float* T, double* pose7);
void pgo_pose_pose7_to_T(const double* pose7, float* T);
int pgo_pose_optimization(const float* Tcw_in, const float* kp_xy, const int32_t* kp_octave, const float* mp_xyz,
                          const uint8_t* has_map_point, int n, const float* inv_level_sigma2, float fx, float fy, float cx,
                          float cy, float* Tcw_out, uint8_t* outlier, uint8_t* round_outliers);
int pgo_match_consecutive(const pgb_keypoint* prev_kps, const uint8_t* prev_desc, int n_prev,
                          const pgb_keypoint* cur_kps, const uint8_t* cur_desc, int n_cur, float flow_x, float flow_y,
                          float maxX, float maxY, float th, const float* scale_factors, int nlevels,
                          int32_t* match_of_cur);
typedef struct pgo_calib pgo_calib;
pgo_calib* pgo_calib_create(const double* gps_v, const int64_t* gps_t, int n_gps, const double* gyro_xyz,
                            const int64_t* gyro_t, int64_t n_gyro, const double* acc_xyz, const int64_t* acc_t,
                            int64_t n_acc);
void pgo_calib_destroy(pgo_calib*);
int64_t pgo_calib_merged_count(const pgo_calib*);
void pgo_calib_merged_events(const pgo_calib*, int64_t* t, int64_t* gi, int64_t* ai);
int64_t pgo_calib_num_intervals(const pgo_calib*);
void pgo_calib_intervals(const pgo_calib*, int64_t* ref_idx, int64_t* merged_idx, int64_t* start, int64_t* end);
double pgo_calib_eval(const pgo_calib*, const double* x, double* grad);
double pgo_calib_eval_core(pgo_calib*, const double* x, double* grad);
int pgo_calib_minimize(pgo_calib*, double* x, double* fx, int max_iterations, double epsilon, int* n_eval);
int pgo_calib_minimize_core(pgo_calib*, double* x, double* fx, int max_iterations, double epsilon, int* n_eval);
int pgo_calib_minimize_literal_driver_core_eval(pgo_calib*, double* x, double* fx, int max_iterations, double epsilon,
                                                int* n_eval);
int64_t pgo_calib_integrate(const pgo_calib*, const double* x, int64_t cap, int64_t* idx, double* speed, double* quat,
                            double* vel, int64_t* dur);
int64_t pgo_calib_integrate_core(pgo_calib*, const double* x, int64_t cap, int64_t* idx, double* speed, double* vel,
                                 int64_t* dur);
void pgo_smooth_time_series(const double* values, const double* times, int64_t n, const double* target, int64_t nt,
                            double sigma, double* out);
int64_t pgo_fit_motion(const double* gps_v, const int64_t* gps_t, int n_gps, const double* gyro_xyz, const int64_t* gyro_t,
                       int64_t n_gyro, const double* acc_xyz, const int64_t* acc_t, int64_t n_acc, int batch_size,
                       int shift_step, int max_iters, double sigma, int mode, int64_t cap, int64_t* out_idx,
                       int64_t* out_t_usec, double* out_avg, double* out_smoothed, double* x_out, int32_t* iters_out,
                       double* fx_out, int64_t* n_evals_total);
int pgo_forward_axis_sum(const double* gps_v, const int64_t* gps_t, int n_gps, const double* gyro_xyz,
                         const int64_t* gyro_t, int64_t n_gyro, const double* acc_xyz, const int64_t* acc_t,
                         int64_t n_acc, int batch_size, int shift_step, const double* x_all, int mode, double min_vel,
                         double min_rot, double* sum_out, int32_t* windows_used);
void pgo_pca3(const double* rows, int64_t n, double* eigvec9, double* eigval3, double* mean3);
int64_t pgo_principal_rotation_axes(const double* gyro_xyz, const int64_t* gyro_t, int64_t n, int64_t interval_usec,
                                    double* axes9, double* rows_out, int64_t cap);
void pgo_angular_velocities_around_axis(const double* gyro_xyz, int64_t n, const double* axis, double* out);
int pgo_time_averaged_values(const double* values, const int64_t* times, int64_t n, const int64_t* ft, int64_t n_frames,
                             double* out, uint8_t* valid);
double pgo_bench_extract_match(const uint8_t* frames, int n, int w, int h, const float* flow_xy, int nfeatures,
                               float scale, int nlevels, int iniTh, int minTh, float th, int nthreads,
                               int64_t* total_kps, int64_t* total_matches, int32_t* kps_per_frame,
                               int32_t* matches_per_frame);
#ifdef __cplusplus
}
#endif
#endif
