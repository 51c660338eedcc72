// NcclComposite.h -- single-process NCCL communicator over the active devices and one float-sum reduce to rank 0.
// Kept apart from the host classes because <cuda_runtime.h> and the host's device-layout aliases (float3, int2...)
// cannot share a translation unit.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

struct NcclGroup;
// ordinals: CUDA device ordinals, rank i = ordinals[i].  Throws std::runtime_error on failure.
NcclGroup* ncclGroupCreate(int count, const int* ordinals);
void ncclGroupDestroy(NcclGroup* group);
// sum of src[i] (on rank i, `count` floats each) into dstRoot on rank 0; stream[i] = cudaStream_t value of rank i.
void ncclGroupReduceSum(NcclGroup* group, const uint64_t* src, uint64_t dstRoot, size_t count, const uint64_t* streams);

// ---- one process per GPU (torchrun-style deployment) ----
// rank 0 creates the id, the launcher distributes the 128 bytes out of band (torch.distributed broadcast, a file, MPI...).
void ncclProcessUniqueId(char out[128]);
// Joins rank `rank` of `world` on CUDA device `ordinal`.  Collective over all ranks.  Throws std::runtime_error.
NcclGroup* ncclProcessGroupJoin(int rank, int world, const char id[128], int ordinal);
// mean over the ranks of `count` floats at src into dst on rank 0 (dst may be 0 elsewhere), on `stream`.  Collective.
void ncclProcessGroupReduceMean(NcclGroup* group, uint64_t src, uint64_t dst, size_t count, uint64_t stream);
