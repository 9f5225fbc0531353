// abi.cu -- version / error-text entry points of the C ABI (include/sparenet_b200.h).
#include "common.cuh"

SNB_API int snb_version(void) { return 200; }  // 0.2.0: round 2 (tcgen05 GEMM, tails, thin / linear / kNN kernels, multi-tensor copy)

SNB_API const char* snb_strerror(int code) {
  switch (code) {
    case SNB_OK: return "ok";
    case SNB_EINVAL: return "invalid argument";
    case SNB_ELIMIT: return "shape outside the documented limits of this op";
    case SNB_EWORKSPACE: return "workspace missing or too small";
    case SNB_EALIGN: return "pointer not 16-byte aligned";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown sparenet_b200 error";
}
