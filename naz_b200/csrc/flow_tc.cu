// tcgen05 engine — placeholder until the tensor-core kernels land.
#include "nazb_internal.h"
bool nazb_tc_supported(const FlowGeom&, std::string* why) { if (why) *why = "tcgen05 engine not built yet"; return false; }
cudaError_t nazb_tc_create(nazb_handle*) { return cudaErrorNotSupported; }
void nazb_tc_destroy(nazb_handle*) {}
cudaError_t nazb_tc_pack(nazb_handle*, const float* const*, const float* const*, const int64_t*, const int64_t*,
                         const float* const*, const float*, float, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t nazb_tc_launch(const nazb_handle*, const IoArgs&, int, cudaStream_t) { return cudaErrorNotSupported; }
int64_t nazb_tc_packed_bytes(const nazb_handle*) { return 0; }
