/* ORACLE shim (test infrastructure): API-compatible stand-in for ggplib's statemachine/basestate.h,
 * covering exactly the surface the reference's search core uses:
 * gdltransformer.h:24,48, gdltransformer.cpp:56-62, puct/node.cpp:51,103-105, puct/evaluator.h:143,
 * uniquestates.h:24,35-50. Bit layout: bit i -> data word i/64, bit i%64 (LSB first), identical in
 * memory to the packed uint32 rows used by libgz_b200 on a little-endian host. */
#pragma once

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

namespace GGPLib {

struct BaseState {
    typedef uint64_t ArrayType;

    int size;        // number of bases (bits)
    int byte_count;  // bytes in data[] (multiple of 16)
    ArrayType data[0];

    static int byteCountFor(int size) { return ((size + 127) / 128) * 16; }
    static size_t mallocSize(int size) { return sizeof(BaseState) + byteCountFor(size); }

    void init(int sz) {
        this->size = sz;
        this->byte_count = byteCountFor(sz);
        std::memset(this->data, 0, this->byte_count);
    }
    bool get(int i) const { return (data[i >> 6] >> (i & 63)) & 1u; }
    void set(int i, bool v) {
        if (v) data[i >> 6] |= (ArrayType(1) << (i & 63));
        else data[i >> 6] &= ~(ArrayType(1) << (i & 63));
    }
    void assign(const BaseState* o) { std::memcpy(data, o->data, byte_count); }
    bool equals(const BaseState* o) const { return std::memcmp(data, o->data, byte_count) == 0; }

    struct Hasher {
        const ArrayType* mask;
        size_t operator()(const BaseState* bs) const {
            uint64_t h = 0xcbf29ce484222325ULL;
            const int n = bs->byte_count / 8;
            for (int i = 0; i < n; i++) {
                uint64_t w = bs->data[i];
                if (mask) w &= mask[i];
                h ^= w + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
            }
            return (size_t)h;
        }
    };
    struct Equals {
        const ArrayType* mask;
        bool operator()(const BaseState* a, const BaseState* b) const {
            const int n = a->byte_count / 8;
            for (int i = 0; i < n; i++) {
                uint64_t x = a->data[i], y = b->data[i];
                if (mask) { x &= mask[i]; y &= mask[i]; }
                if (x != y) return false;
            }
            return true;
        }
    };

    template <typename V>
    using HashMapMasked = std::unordered_map<const BaseState*, V, Hasher, Equals>;
    template <typename V>
    using HashMap = std::unordered_map<const BaseState*, V, Hasher, Equals>;

    template <typename V>
    static HashMapMasked<V>* makeMaskedMap(ArrayType* mask) {
        return new HashMapMasked<V>(16, Hasher{mask}, Equals{mask});
    }
};

}  // namespace GGPLib
