// shuffle.cu — row -> rank radix partition (placeholder until the kernel lands; fails loudly).
#include "common.cuh"
extern "C" {
int b200_hash_to_rank(const b200_table*, int32_t, int32_t*, void*) { b200::set_last_error("b200_hash_to_rank: not implemented yet"); return -1; }
int b200_shuffle_partition(const b200_table*, int64_t, int32_t, b200_table*, int64_t*, void*) { b200::set_last_error("b200_shuffle_partition: not implemented yet"); return -1; }
}
