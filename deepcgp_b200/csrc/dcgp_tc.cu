// placeholder until the tcgen05 path lands
#include "dcgp_tc.cuh"
namespace dcgp {
void tc_carve_prep(TcPrep& t, int M, int Mp, int R, int L, void* buf) { memset(&t, 0, sizeof(t)); t.bytes = 0; }
int tc_pack_operands(const TcPrep&, const double*, int, const double*, const double*, int, int, int, cudaStream_t) { set_error("tc path not built"); return DCGP_ERR_ARG; }
int tc_pack_z(const TcPrep&, const double*, int, int, double, cudaStream_t) { set_error("tc path not built"); return DCGP_ERR_ARG; }
void tc_carve_cond(TcCondWork& w, int, int, int, size_t, void*) { memset(&w, 0, sizeof(w)); }
int tc_split_rows(const float*, int, int, const TcCondWork&, cudaStream_t) { set_error("tc path not built"); return DCGP_ERR_ARG; }
int tc_cond(const TcPrep&, const TcCondWork&, int, int, int, float*, float*, cudaStream_t) { set_error("tc path not built"); return DCGP_ERR_ARG; }
void tc_carve_apply(TcApplyWork& a, int, int, int, int, int, size_t, size_t, void*) { memset(&a, 0, sizeof(a)); }
int tc_layer_apply(const dcgp_layer_desc*, const View&, const TcPrep&, const TcApplyWork&, const double*, const float*, int, float*, float*, float*, float*, cudaStream_t) { set_error("tc path not built"); return DCGP_ERR_ARG; }
}
