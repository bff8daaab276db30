// nccl.h (shim) -- TEST INFRASTRUCTURE: in-process stand-in for the few NCCL
// entry points halo.cu uses (ranks are threads of one process; send/recv go
// through a mailbox).  See cuda_runtime.h in this directory.
#pragma once
#include <stddef.h>
#include "cuda_runtime.h"
typedef enum { ncclSuccess = 0, ncclInternalError = 3, ncclInvalidArgument = 4 } ncclResult_t;
typedef enum { ncclDouble = 8 } ncclDataType_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef struct emuNcclComm* ncclComm_t;
extern "C" {
ncclResult_t ncclGetUniqueId(ncclUniqueId*);
ncclResult_t ncclCommInitRank(ncclComm_t*, int, ncclUniqueId, int);
ncclResult_t ncclCommDestroy(ncclComm_t);
ncclResult_t ncclSend(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
ncclResult_t ncclRecv(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
ncclResult_t ncclGroupStart();
ncclResult_t ncclGroupEnd();
const char* ncclGetErrorString(ncclResult_t);
}
