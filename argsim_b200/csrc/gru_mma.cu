// placeholder until the persistent recurrence lands
#include "kernels.h"
#include "plan.h"
struct GruMmaCtx { int dev; };
GruMmaCtx* gru_mma_create(int device) { return new GruMmaCtx{device}; }
void gru_mma_destroy(GruMmaCtx* c) { delete c; }
bool gru_mma_supported(int) { return false; }
void gru_mma_fwd(GruMmaCtx*, const GruFwdArgs*, int, const SeqPlan&, const int*, const int*, int, cudaStream_t) { throw std::runtime_error("gru_mma not built"); }
void gru_mma_bwd(GruMmaCtx*, const GruBwdArgs*, int, const SeqPlan&, const int*, const int*, int, cudaStream_t) { throw std::runtime_error("gru_mma not built"); }
