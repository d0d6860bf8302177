// stubs.cu — placeholders replaced by ivf_pq.cu / hnsw.cu as they land.
#include "index.h"
namespace b200vs {
#ifndef B200VS_HAVE_PQ
IndexBase* make_ivf_pq(b200vs_metric, int, const b200vs_params&) { fail(B200VS_EVECTOR_NOT_SUPPORT, "IVF_PQ not built yet"); }
#endif
#ifndef B200VS_HAVE_HNSW
IndexBase* make_hnsw(b200vs_metric, int, const b200vs_params&) { fail(B200VS_EVECTOR_NOT_SUPPORT, "HNSW not built yet"); }
#endif
}
