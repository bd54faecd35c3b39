// dense_dmma.cu -- fused FP64 tensor-core (DMMA) statistics kernel of the dense E-step (placeholder:
// the generic path is used until the fused kernel lands).
#include "ppca.cuh"
namespace gplvm {
bool dense_dmma_eligible(const gplvm_ppca_session*) { return false; }
int dense_dmma_prepare(gplvm_ppca_session*) { return GPLVM_OK; }
int dense_dmma_stats(gplvm_ppca_session*, ShardState&, const double*, const double*) { return fail(GPLVM_ERR_STATE, "dmma path not built"); }
int dense_dmma_release(gplvm_ppca_session*) { return GPLVM_OK; }
int dense_dmma_nparts(const gplvm_ppca_session*, const ShardState&) { return 0; }
}
