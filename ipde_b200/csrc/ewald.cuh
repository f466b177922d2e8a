#pragma once
#include "grid.cuh"

struct EwaldPlan;
// All tile/table arrays are host arrays (copied); source and target arrays are device arrays (copied).
int ewald_plan_create(int kind, double kappa, double xi, double h, double rc, int m, double x0, double y0, int px, int py,
                      double L, double c0, double c1, int64_t nsrc, const double* sx_dev, const double* sy_dev, int nint,
                      int ndeg, int nlog, const double* tabS_host, const double* tabA_host, int ntiles,
                      const int* tile_ix_host, const int* tile_iy_host, const int* tile_ptr_host,
                      const int* tile_src_host, const int* tile_map_host, EwaldPlan** out);
void ewald_plan_destroy(EwaldPlan* p);
int ewald_set_grid_targets(EwaldPlan* p, int64_t n, const int64_t* lin_dev, const int64_t* loc_dev, const int64_t* out_pos_dev);
int ewald_set_point_targets(EwaldPlan* p, int64_t n, const double* tx_dev, const double* ty_dev);
int ewald_apply(EwaldPlan* p, const double* q, double* grid_out, double* point_out, int accumulate_points, cudaStream_t st);
int ewald_debug_buffer(EwaldPlan* p, int which, double** ptr, int64_t* n);
int ewald_set_gradient_tables(EwaldPlan* p, const double* dS_host, const double* dA_host);
int ewald_apply_fields(EwaldPlan* p, const double* q, double xc, double yc, const double* gw, double* const* gdst,
                       const double* pw, double* const* pdst, cudaStream_t st);
