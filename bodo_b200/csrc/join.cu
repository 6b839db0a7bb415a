// join.cu — streaming hash join (placeholder until the kernels land; fails loudly).
#include "common.cuh"
extern "C" {
void* b200_join_state_init(int64_t, const int8_t*, const int8_t*, int32_t, const int8_t*, const int8_t*, int32_t, uint64_t, int32_t, int32_t, int64_t, int32_t, int64_t, void*) { b200::set_last_error("b200 join: not implemented yet"); return nullptr; }
int b200_join_build_consume_batch(void*, const b200_table*, int32_t, int32_t*) { b200::set_last_error("b200 join: not implemented yet"); return -1; }
int b200_join_probe_consume_batch(void*, const b200_table*, const uint64_t*, int64_t, const uint64_t*, int64_t, b200_table*, int64_t*, int32_t, int32_t*) { b200::set_last_error("b200 join: not implemented yet"); return -1; }
void b200_delete_join_state(void*) {}
int64_t b200_join_get_metric(void*, int32_t) { return -1; }
}
