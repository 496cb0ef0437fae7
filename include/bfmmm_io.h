/*
 * bfmmm_io.h -- the reference's stored-sample file layout.
 *
 * BFMMM_MTT_warm_start writes every r_stored_iters iterations (BFMMM.h:1680-1746; cov-adj
 * :5152-5168) the thinned batch as  directory + {Nu,Chi,Pi,alpha_3,A,Delta,Sigma,Tau,Z}{q}.txt
 * (arma_ascii: ARMA_MAT_TXT_FN008 / ARMA_CUB_TXT_FN008) and {Gamma,Phi}{q}.txt (field<cube>,
 * arma_binary: ARMA_FLD_BIN of ARMA_CUB_BIN_FN008 elements); ReadVec/ReadMat/ReadCube/ReadFieldCube
 * (src/UserFunctions.cpp:2158-2355) and every src/PostProcessing.cpp function read them back.
 * These entry points write and read exactly those bytes, so chains produced by this engine are
 * consumed by the unchanged R post-processing.  Data are column-major doubles.
 */
#ifndef BFMMM_IO_H
#define BFMMM_IO_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { BFMMM_FILE_MAT_TXT = 1, BFMMM_FILE_CUBE_TXT = 2, BFMMM_FILE_FIELD_CUBE_BIN = 3, BFMMM_FILE_MAT_BIN = 4,
       BFMMM_FILE_CUBE_BIN = 5, BFMMM_FILE_FIELD_MAT_BIN = 6 };

int bfmmm_save_mat_txt(const char* path, const double* data, int64_t n_rows, int64_t n_cols);
int bfmmm_save_cube_txt(const char* path, const double* data, int64_t n_rows, int64_t n_cols, int64_t n_slices);
/* field of n_field_rows x n_field_cols cubes (all r x c x s), elements stored back to back in the
 * field's column-major order, each cube column-major */
int bfmmm_save_field_cube_bin(const char* path, const double* data, int64_t n_field_rows, int64_t n_field_cols,
                              int64_t r, int64_t c, int64_t s);
/* kind and dims = {n_rows, n_cols, n_slices, n_field_rows, n_field_cols} of a stored file */
int bfmmm_file_info(const char* path, int32_t* kind, int64_t* dims);
/* reads any of the formats above into `out` (capacity in doubles); fields are concatenated */
int bfmmm_load(const char* path, double* out, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif
