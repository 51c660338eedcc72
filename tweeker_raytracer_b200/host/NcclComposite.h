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
