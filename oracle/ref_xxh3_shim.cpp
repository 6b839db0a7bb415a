// Shim that exposes the REFERENCE's own hash_inner_32 (bodo/libs/vendored/_murmurhash3.h:59-68:
// low 32 bits of XXH3_64bits_withSeed) through a C ABI. Compiled against the vendored header in
// place under /root/reference (never copied); output goes to oracle/_ref/. Test infrastructure only.
#define XXH_INLINE_ALL
#include "xxhash.h"
#include <cstdint>
extern "C" uint32_t ref_hash_inner_32_i64(int64_t v, uint32_t seed) {
    return static_cast<uint32_t>(XXH3_64bits_withSeed((const void*)&v, sizeof(v), seed));
}
extern "C" uint32_t ref_hash_inner_32_i32(int32_t v, uint32_t seed) {
    return static_cast<uint32_t>(XXH3_64bits_withSeed((const void*)&v, sizeof(v), seed));
}
extern "C" uint64_t ref_xxh3_64_i64(int64_t v, uint32_t seed) {
    return XXH3_64bits_withSeed((const void*)&v, sizeof(v), seed);
}
