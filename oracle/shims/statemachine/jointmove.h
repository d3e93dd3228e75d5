/* ORACLE shim for ggplib's statemachine/jointmove.h (used inline at the tail of PuctNodeChild,
 * src/cpp/puct/node.h:42-57, node.cpp:55,140-141). */
#pragma once

#include <cstdint>
#include <cstdlib>

namespace GGPLib {

struct JointMove {
    typedef uint16_t IndexType;
    int count;
    IndexType indices[0];

    static size_t mallocSize(int role_count) { return sizeof(JointMove) + role_count * sizeof(IndexType); }
    void setSize(int n) { count = n; }
    int getCount() const { return count; }
    IndexType get(int i) const { return indices[i]; }
    void set(int i, IndexType v) { indices[i] = v; }
    void assign(const JointMove* o) { for (int i = 0; i < count; i++) indices[i] = o->indices[i]; }
    bool equals(const JointMove* o) const {
        for (int i = 0; i < count; i++) if (indices[i] != o->indices[i]) return false;
        return true;
    }
};

}  // namespace GGPLib
