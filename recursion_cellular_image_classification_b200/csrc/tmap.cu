// tmap.cu — host-side CUtensorMap encoding (driver entry point fetched through the runtime).
#include "common.cuh"
#include "ptx.cuh"

namespace rxb {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, void* base, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return set_error(RXB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, dt, (cuuint32_t)rank, base, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(RXB_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d): rank %d base %p dims [%llu %llu %llu %llu] "
                     "strides [%llu %llu %llu] box [%u %u %u %u] swz %d",
                     (int)r, rank, base, (unsigned long long)dims[0],
                     (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                     (unsigned long long)(rank > 3 ? dims[3] : 0),
                     (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
                     (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
                     (unsigned long long)(rank > 3 ? strides_bytes[2] : 0), box[0], rank > 1 ? box[1] : 0,
                     rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, (int)swz);
  }
  return RXB_OK;
}

}  // namespace rxb
