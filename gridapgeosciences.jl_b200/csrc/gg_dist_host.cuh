// Host-side structures of the partitioned-model path (gg_dist.cu).
#pragma once
#include <map>
#include <memory>
#include <vector>
#include "gg_common.cuh"
#include "gg_dist.cuh"

namespace gg {
struct SpacePlan {
  int kind = 0, order = 0, mloc = 0;
  int64_t n_global = 0, n_owned = 0, slot_capacity = 0;
  std::vector<int64_t> l2g;         // local -> global DOF (0-based, ascending)
  std::vector<int32_t> owner;       // owner rank of every local DOF
  std::vector<int32_t> cell_dofs;   // [n_local_cells][mloc] local ids
  std::vector<int8_t> signs;        // RT: signs of the GLOBAL numbering
  std::vector<int32_t> cob_master_panel;   // vector H1: master panel / chart point of every local (cell, node)
  std::vector<double> cob_master_point;
  struct Block { int64_t start, count; int size; };
  std::vector<Block> blocks;
  std::vector<int32_t> peers, send_ptr, send_idx, send_remote, recv_ptr, recv_idx;
};
SpacePlan& plan_space(gg_plan* P, int kind, int order);
DistDev dist_dev(gg_space* s);
void vec_consistent_async(gg_space* s, double* vec);
// throws GG_ERR_COMM if a wait on this communicator's window has timed out (the stream must be idle); no-op for nullptr
void comm_check(gg_comm* c);
}  // namespace gg

struct gg_plan {
  int n = 0, nparts = 1, part = 0;
  double radius = 1.0;
  std::unique_ptr<gg_mesh> gmesh;              // global mesh, host arrays only
  // 0: the reference's linear partition of the panel-major cell order (src/Distributed/AtlasDistributedDiscreteModels.jl:31);
  // 1: panel strips -- rank r owns the same range of in-panel cell indices on all six panels, so that every rank keeps the six
  //    congruent cells of a geometry class together (opt-in; the partition does not change any global result)
  int mode = 0;
  std::vector<int32_t> cell_owner;             // owner rank of every global cell
  std::vector<int64_t> n_owned;                // owned cells of every rank
  std::vector<int64_t> first, last;            // smallest owned cell / one past the largest (the owned RANGE when mode == 0)
  std::vector<std::vector<int64_t>> cells;     // local cells (owned + ghost, ascending global id) of every rank
  std::map<int, std::unique_ptr<gg::SpacePlan>> spaces;
};

struct gg_comm {
  gg_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  int64_t slot_doubles = 0;
  size_t bytes = 0;
  void* window = nullptr;           // this rank's window: WinHeader + exchange slot
  std::vector<void*> peer;          // [world] device pointers (IPC mappings for the other ranks)
  gg::DevBuf<void*> d_win, d_slot;  // the same pointers on the device
  bool connected = false;
};
