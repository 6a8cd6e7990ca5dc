// temporary stubs for entry points that are not built yet
#include "lz_common.cuh"
#define STUB(name) { lz_set_error(name ": not built yet"); return LZ_ERR_UNSUPPORTED; }
extern "C" {
int lz_ritz(int, int, const double *, const double *, const double *, int, double *, double *) STUB("lz_ritz")
int lz_comm_unique_id(void *) STUB("lz_comm_unique_id")
int lz_comm_init(lz_ctx *, int, int, const void *) STUB("lz_comm_init")
int lz_comm_destroy(lz_ctx *) { return LZ_OK; }
int lz_partition_rows(int64_t, int, int, int64_t *, int64_t *) STUB("lz_partition_rows")
int lz_gen_laplacian3d_shard(lz_ctx *, int64_t, int64_t, int64_t, int, int, lz_matrix **) STUB("lz_gen_laplacian3d_shard")
int lz_gen_laplacian2d_shard(lz_ctx *, int64_t, int64_t, int, int, lz_matrix **) STUB("lz_gen_laplacian2d_shard")
int lz_vector_lanczos_sharded(lz_ctx *, const lz_matrix *, const double *, int, int, double *, double *) STUB("lz_vector_lanczos_sharded")
}
