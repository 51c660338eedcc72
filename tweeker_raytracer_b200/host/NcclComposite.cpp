#include "NcclComposite.h"

#include <cstring>
#include <stdexcept>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>

struct NcclGroup
{
  std::vector<ncclComm_t> comms;
  std::vector<int> ordinals;
};

static void check(ncclResult_t r, const char* what)
{
  if (r != ncclSuccess) throw std::runtime_error(std::string("ERROR: ") + what + ": " + ncclGetErrorString(r));
}

NcclGroup* ncclGroupCreate(int count, const int* ordinals)
{
  NcclGroup* g = new NcclGroup();
  g->ordinals.assign(ordinals, ordinals + count);
  g->comms.resize((size_t)count);
  try { check(ncclCommInitAll(g->comms.data(), count, g->ordinals.data()), "ncclCommInitAll"); }
  catch (...) { delete g; throw; }
  return g;
}

void ncclGroupDestroy(NcclGroup* group)
{
  if (!group) return;
  for (ncclComm_t c : group->comms) ncclCommDestroy(c);
  delete group;
}

void ncclGroupReduceSum(NcclGroup* group, const uint64_t* src, uint64_t dstRoot, size_t count, const uint64_t* streams)
{
  check(ncclGroupStart(), "ncclGroupStart");
  for (size_t i = 0; i < group->comms.size(); ++i)
  {
    cudaSetDevice(group->ordinals[i]);
    check(ncclReduce((const void*)(uintptr_t)src[i], (void*)(uintptr_t)dstRoot, count, ncclFloat, ncclSum, 0, group->comms[i],
                     (cudaStream_t)(uintptr_t)streams[i]), "ncclReduce");
  }
  check(ncclGroupEnd(), "ncclGroupEnd");
}

void ncclProcessUniqueId(char out[128])
{
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  check(ncclGetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out, &id, sizeof(id));
}

NcclGroup* ncclProcessGroupJoin(int rank, int world, const char idBytes[128], int ordinal)
{
  NcclGroup* g = new NcclGroup();
  g->ordinals.assign(1, ordinal);
  g->comms.resize(1);
  ncclUniqueId id;
  std::memcpy(&id, idBytes, sizeof(id));
  cudaSetDevice(ordinal);
  try { check(ncclCommInitRank(&g->comms[0], world, id, rank), "ncclCommInitRank"); }
  catch (...) { delete g; throw; }
  return g;
}

void ncclProcessGroupReduceMean(NcclGroup* group, uint64_t src, uint64_t dst, size_t count, uint64_t stream)
{
  cudaSetDevice(group->ordinals[0]);
  check(ncclReduce((const void*)(uintptr_t)src, (void*)(uintptr_t)dst, count, ncclFloat, ncclAvg, 0, group->comms[0],
                   (cudaStream_t)(uintptr_t)stream), "ncclReduce(avg)");
}
