// tiled.cu -- node-tile owner-computes assembly (CGASM_SCATTER_TILED). Placeholder until the
// tiled kernels land: selecting the variant reports CGASM_EUNSUPPORTED.
#include "cgasm_internal.h"

namespace cgasm {
struct TilePlan {};
int tiles_build(Handle*) { CG_FAIL(CGASM_EUNSUPPORTED, "tiled scatter not built yet"); }
void tiles_free(Handle* h) {
  delete h->tiles;
  h->tiles = nullptr;
}
int tiles_momentum(Handle*, const MomentumArgs&, bool, bool) { CG_FAIL(CGASM_EUNSUPPORTED, "tiled scatter not built yet"); }
int tiles_advdiff(Handle*, const AdvDiffArgs&) { CG_FAIL(CGASM_EUNSUPPORTED, "tiled scatter not built yet"); }
}  // namespace cgasm
