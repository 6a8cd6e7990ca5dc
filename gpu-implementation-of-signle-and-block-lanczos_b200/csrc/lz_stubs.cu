// temporary stubs for entry points that are not built yet
#include "lz_common.cuh"
#define STUB(name) { lz_set_error(name ": not built yet"); return LZ_ERR_UNSUPPORTED; }
extern "C" {
int lz_spmm(lz_ctx *, const lz_matrix *, int, const double *, int64_t, double *, int64_t) STUB("lz_spmm")
int lz_mm_tt(lz_ctx *, int64_t, int, const double *, int64_t, double *) STUB("lz_mm_tt")
int lz_mm_tt2(lz_ctx *, int64_t, int, const double *, int64_t, const double *, int64_t, double *) STUB("lz_mm_tt2")
int lz_mm_ts(lz_ctx *, int64_t, int, double, double, const double *, int64_t, const double *, double *, int64_t) STUB("lz_mm_ts")
int lz_sqrtm(lz_ctx *, int, double *, double *) STUB("lz_sqrtm")
int lz_copy_row(lz_ctx *, int64_t, int, const double *, int64_t, double *, int64_t) STUB("lz_copy_row")
int lz_assemble_T(lz_ctx *, int, int, const double *, const double *, double *) STUB("lz_assemble_T")
int lz_block_lanczos(lz_ctx *, const lz_matrix *, const double *, int64_t, int, int, int64_t, int, double *, double *, double *) STUB("lz_block_lanczos")
int lz_ritz(int, int, const double *, const double *, const double *, int, double *, double *) STUB("lz_ritz")
int lz_comm_unique_id(void *) STUB("lz_comm_unique_id")
int lz_comm_init(lz_ctx *, int, int, const void *) STUB("lz_comm_init")
int lz_comm_destroy(lz_ctx *) { return LZ_OK; }
int lz_partition_rows(int64_t, int, int, int64_t *, int64_t *) STUB("lz_partition_rows")
int lz_gen_laplacian3d_shard(lz_ctx *, int64_t, int64_t, int64_t, int, int, lz_matrix **) STUB("lz_gen_laplacian3d_shard")
int lz_gen_laplacian2d_shard(lz_ctx *, int64_t, int64_t, int, int, lz_matrix **) STUB("lz_gen_laplacian2d_shard")
int lz_vector_lanczos_sharded(lz_ctx *, const lz_matrix *, const double *, int, int, double *, double *) STUB("lz_vector_lanczos_sharded")
}
