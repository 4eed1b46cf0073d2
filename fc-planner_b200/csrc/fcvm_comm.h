// fcvm_comm.h — the two collectives of the sharded manager, behind one function pointer:
//   FCVM_COMM_ALLGATHER       in place over `world` equal slices of n bytes (this rank's slice already filled)
//   FCVM_COMM_ALLREDUCE_MAX   in place, element-wise MAX over n int64 values
// Either the caller supplies the function (fcvm_set_shard: e.g. torch.distributed from Python) or the library talks to
// NCCL itself (fcvm_set_shard_nccl): libnccl.so.2 is dlopen'ed on first use, so libfcvm.so carries no link-time dependency
// on it and single-GPU users never load it.  All collectives are enqueued on the caller's stream (NVLink 5 / NVSwitch).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/fcvm.h"

namespace fcvm {

struct NcclComm;   // opaque: a communicator created by nccl_comm_create

// 0 / FCVM_E* (message via fail()).  id128 = the 128 bytes of an ncclUniqueId.
int nccl_unique_id(uint8_t* id128);
int nccl_comm_create(const uint8_t* id128, int rank, int world, NcclComm** out);   // the CUDA device must be current
void nccl_comm_destroy(NcclComm* c);
// fcvm_comm_fn-shaped entry: user = NcclComm*
int nccl_comm_call(void* user, int32_t op, void* buf_dev, int64_t n, void* stream);
const char* nccl_version_string();

}  // namespace fcvm
