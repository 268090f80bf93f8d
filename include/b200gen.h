/*
 * b200gen.h -- C ABI of libb200gen.so: synthetic-workload generators on the device.  BENCH / TEST PLUMBING, not part of
 * the product library (libb200spmv.so exports none of this).  Bit-identical to the CPU generators of the checker
 * (oracle/spmv_oracle.c), which restate cuSPARSE/cg/cg_example.c:71-128, cuSPARSE/bicgstab/bicgstab_example.c:69-127 and
 * cuDSS/simple_residual/laplace_generator.hxx:34-107, so full-size benchmark matrices never cross PCIe.
 */
#ifndef B200GEN_H_
#define B200GEN_H_

#include <stddef.h>
#include <stdint.h>

#define B200GEN_EXPORT /* exported from libb200gen.so */

#ifdef __cplusplus
extern "C" {
#endif

B200GEN_EXPORT int b200gen_rmat_keys(void* stream, uint64_t seed, int64_t e0, int64_t count, int32_t scale,
                                      uint64_t tA, uint64_t tAB, uint64_t tABC, int64_t rows, int64_t cols,
                                      int64_t* keys_out);
B200GEN_EXPORT int b200gen_uniform(void* stream, int dtype, uint64_t seed, int64_t i0, int64_t count, void* out);
B200GEN_EXPORT int b200gen_stencil5_counts(void* stream, int32_t grid, int32_t* counts_out);
B200GEN_EXPORT int b200gen_stencil5_fill(void* stream, int32_t grid, double mass, double ux, double uy,
                                          const int32_t* row_offsets, int32_t* col_out, double* val_out);
B200GEN_EXPORT int b200gen_laplace7_counts(void* stream, int32_t nx, int32_t* counts_out);
B200GEN_EXPORT int b200gen_laplace7_fill(void* stream, int dtype, int32_t nx, const int32_t* row_offsets,
                                          int32_t* col_out, void* val_out);


#ifdef __cplusplus
}
#endif
#endif /* B200GEN_H_ */
