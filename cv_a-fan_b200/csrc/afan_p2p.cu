// afan_p2p.cu -- peer-mapped device memory for the fused statistics exchange (one process per GPU).
// Plain cudaMalloc + cudaIpc handles: the host side (cv_a-fan_b200/p2p.py) all-gathers the 64-byte handles over
// torch.distributed and opens every peer's mailbox; NVLink/NVSwitch P2P access is enabled lazily by the open.
#include <string.h>

#include "afan_common.cuh"

static int cuda_rc(cudaError_t e) {
    if (e == cudaSuccess) return AFAN_OK;
    cudaGetLastError();
    return AFAN_ERR_LAUNCH;
}

AFAN_EXPORT int afan_p2p_alloc(void** ptr, int64_t bytes) {
    if (!ptr) return AFAN_ERR_NULL;
    if (bytes <= 0) return AFAN_ERR_SIZE;
    int rc = cuda_rc(cudaMalloc(ptr, static_cast<size_t>(bytes)));
    if (rc != AFAN_OK) return rc;
    return cuda_rc(cudaMemset(*ptr, 0, static_cast<size_t>(bytes)));
}

AFAN_EXPORT int afan_p2p_free(void* ptr) { return ptr ? cuda_rc(cudaFree(ptr)) : AFAN_OK; }

AFAN_EXPORT int afan_p2p_get_handle(void* ptr, void* handle_out_64) {
    if (!ptr || !handle_out_64) return AFAN_ERR_NULL;
    cudaIpcMemHandle_t h;
    int rc = cuda_rc(cudaIpcGetMemHandle(&h, ptr));
    if (rc != AFAN_OK) return rc;
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle_out_64, &h, 64);
    return AFAN_OK;
}

AFAN_EXPORT int afan_p2p_open_handle(const void* handle_64, void** ptr_out) {
    if (!handle_64 || !ptr_out) return AFAN_ERR_NULL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_64, 64);
    return cuda_rc(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
}

AFAN_EXPORT int afan_p2p_close_handle(void* ptr) { return ptr ? cuda_rc(cudaIpcCloseMemHandle(ptr)) : AFAN_OK; }
