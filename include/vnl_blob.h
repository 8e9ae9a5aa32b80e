/*
 * vnl_blob.h -- tiny accessors for the flat model / task blobs of vnl_b200.h.
 * Usable from C, C++ and CUDA device code (everything is a word offset lookup).
 */
#ifndef VNL_BLOB_H_
#define VNL_BLOB_H_

#include "vnl_b200.h"

#if defined(__CUDACC__)
#define VNL_HD __host__ __device__ __forceinline__
#else
#define VNL_HD static inline
#endif

VNL_HD int vnl_hdr_i(const uint32_t* blob, int slot) { return (int)blob[slot]; }
VNL_HD float vnl_hdr_f(const uint32_t* blob, int slot) {
  union { uint32_t u; float f; } c;
  c.u = blob[slot];
  return c.f;
}
VNL_HD const int* vnl_field_i(const uint32_t* blob, int f) {
  return (const int*)(blob + blob[VNL_TABLE_OFF + 2 * f]);
}
VNL_HD const float* vnl_field_f(const uint32_t* blob, int f) {
  return (const float*)(blob + blob[VNL_TABLE_OFF + 2 * f]);
}
VNL_HD int vnl_field_n(const uint32_t* blob, int f) { return (int)blob[VNL_TABLE_OFF + 2 * f + 1]; }

#endif /* VNL_BLOB_H_ */
