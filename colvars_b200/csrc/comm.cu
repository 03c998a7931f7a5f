// comm.cu -- multi-GPU: one process per GPU, group1 rows sharded, group2 replicated; the one
// collective of the path is an ncclAllReduce over [value | grad1 | grad2].  NCCL is loaded with
// dlopen so that the copy already mapped into the process (torch's) is the one used.
#include <dlfcn.h>

#include "engine.h"

using namespace b200cv;

namespace {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::mutex mu;
  bool load(std::string &err)
  {
    std::lock_guard<std::mutex> lock(mu);
    if (lib && AllReduce) return true;
    // picks up the copy already mapped into the process (e.g. torch's) or the system one
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      err = std::string("cannot load NCCL: ") + dlerror();
      return false;
    }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !AllGather) {
      err = "NCCL library lacks required symbols";
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;

}  // namespace

namespace b200cv {

int comm_allreduce(b200cv_cvc *h, cudaStream_t stream)
{
  if (!h->comm) return B200CV_OK;  // b200cv_set_shard: the caller reduces
  ncclResult_t r = g_nccl.AllReduce(h->d_reduce, h->d_reduce, h->reduce_count, ncclDouble,
                                    ncclSum, h->comm, stream);
  if (r != ncclSuccess) {
    return fail(h, B200CV_ERROR,
                std::string("ncclAllReduce: ") +
                    (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
  }
  return B200CV_OK;
}

int comm_allgather(b200cv_cvc *h, double *buf, size_t count_per_rank, cudaStream_t stream)
{
  if (!h->comm) return fail(h, B200CV_BUG_ERROR, "all-gather without a communicator");
  // in place: rank r's contribution already sits at buf + r * count_per_rank
  ncclResult_t r = g_nccl.AllGather(buf + (size_t)h->rank * count_per_rank, buf, count_per_rank,
                                    ncclDouble, h->comm, stream);
  if (r != ncclSuccess) {
    return fail(h, B200CV_ERROR,
                std::string("ncclAllGather: ") +
                    (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
  }
  return B200CV_OK;
}

void comm_destroy(b200cv_cvc *h)
{
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  h->comm = nullptr;
}

}  // namespace b200cv

extern "C" {

int b200cv_comm_unique_id(unsigned char id[128])
{
  std::string err;
  if (!g_nccl.load(err)) return fail(nullptr, B200CV_ERROR, err);
  ncclUniqueId uid;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclResult_t r = g_nccl.GetUniqueId(&uid);
  if (r != ncclSuccess) return fail(nullptr, B200CV_ERROR, "ncclGetUniqueId failed");
  std::memcpy(id, &uid, 128);
  return B200CV_OK;
}

int b200cv_set_shard(b200cv_cvc *h, int rank, int world)
{
  if (!h) return B200CV_BUG_ERROR;
  if (world < 1 || rank < 0 || rank >= world) return fail(h, B200CV_INPUT_ERROR, "bad rank/world");
  h->rank = rank;
  h->world = world;
  h->items_dirty = true;
  h->pl_valid = false;
  return B200CV_OK;
}

int b200cv_comm_init(b200cv_cvc *h, const unsigned char id[128], int rank, int world)
{
  if (!h || !id) return B200CV_BUG_ERROR;
  int rc = b200cv_set_shard(h, rank, world);
  if (rc) return rc;
  std::string err;
  if (!g_nccl.load(err)) return fail(h, B200CV_ERROR, err);
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  ncclUniqueId uid;
  std::memcpy(&uid, id, 128);
  ncclResult_t r = g_nccl.CommInitRank(&h->comm, world, uid, rank);
  if (r != ncclSuccess) {
    return fail(h, B200CV_ERROR,
                std::string("ncclCommInitRank: ") +
                    (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
  }
  return B200CV_OK;
}

}  // extern "C"
