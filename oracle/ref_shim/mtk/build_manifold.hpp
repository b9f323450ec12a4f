/*
 * oracle/ref_shim/mtk/build_manifold.hpp -- TEST INFRASTRUCTURE.  MTK_BUILD_MANIFOLD(name, ((type, id)) ((type, id)) ...)
 * as PoseWithVelocity.hpp:18-23 and OrientationState.hpp:20-26 use it: a struct whose members are SubManifolds in
 * declaration order, DOF = the sum of theirs, boxplus / boxminus member-wise on the matching slice of the tangent
 * vector.  Upstream drives this with Boost.Preprocessor; here the sequence is walked by a pair of
 * mutually recursive-looking macros (the usual A / B trick), once per generated section, and a member's start index is
 * the offset of a char array of its DOF in a layout struct declared first.
 */
#ifndef REF_SHIM_MTK_BUILD_MANIFOLD
#define REF_SHIM_MTK_BUILD_MANIFOLD

#include <cstddef>

#include "src/SubManifold.hpp"

#define MTK_SHIM_CAT(a, b) MTK_SHIM_CAT_(a, b)
#define MTK_SHIM_CAT_(a, b) a##b

/* section 1: layout  ->  char id[type::DOF]; */
#define MTK_SHIM_LAYOUT(type, id) char id[type::DOF];
#define MTK_SHIM_LAYOUT_A(x) MTK_SHIM_LAYOUT x MTK_SHIM_LAYOUT_B
#define MTK_SHIM_LAYOUT_B(x) MTK_SHIM_LAYOUT x MTK_SHIM_LAYOUT_A
#define MTK_SHIM_LAYOUT_A_END
#define MTK_SHIM_LAYOUT_B_END
/* section 2: members */
#define MTK_SHIM_MEMBER(type, id) MTK::shim_at<int(offsetof(mtk_layout_, id))>::sub<type> id;
#define MTK_SHIM_MEMBER_A(x) MTK_SHIM_MEMBER x MTK_SHIM_MEMBER_B
#define MTK_SHIM_MEMBER_B(x) MTK_SHIM_MEMBER x MTK_SHIM_MEMBER_A
#define MTK_SHIM_MEMBER_A_END
#define MTK_SHIM_MEMBER_B_END
/* section 3: boxplus */
#define MTK_SHIM_PLUS(type, id) id.boxplus(mtk_vec_ + int(offsetof(mtk_layout_, id)), mtk_scale_);
#define MTK_SHIM_PLUS_A(x) MTK_SHIM_PLUS x MTK_SHIM_PLUS_B
#define MTK_SHIM_PLUS_B(x) MTK_SHIM_PLUS x MTK_SHIM_PLUS_A
#define MTK_SHIM_PLUS_A_END
#define MTK_SHIM_PLUS_B_END
/* section 4: boxminus */
#define MTK_SHIM_MINUS(type, id) id.boxminus(mtk_res_ + int(offsetof(mtk_layout_, id)), mtk_other_.id);
#define MTK_SHIM_MINUS_A(x) MTK_SHIM_MINUS x MTK_SHIM_MINUS_B
#define MTK_SHIM_MINUS_B(x) MTK_SHIM_MINUS x MTK_SHIM_MINUS_A
#define MTK_SHIM_MINUS_A_END
#define MTK_SHIM_MINUS_B_END

#define MTK_BUILD_MANIFOLD(name, entries)                                                                   \
    struct name##_mtk_layout_t {                                                                            \
        MTK_SHIM_CAT(MTK_SHIM_LAYOUT_A entries, _END)                                                       \
    };                                                                                                      \
    struct name {                                                                                           \
        typedef name self;                                                                                  \
        typedef double scalar;                                                                              \
        typedef name##_mtk_layout_t mtk_layout_;                                                            \
        enum { DOF = sizeof(name##_mtk_layout_t) };                                                         \
        MTK_SHIM_CAT(MTK_SHIM_MEMBER_A entries, _END)                                                       \
        void boxplus(const scalar* mtk_vec_, scalar mtk_scale_ = 1)                                         \
        {                                                                                                   \
            MTK_SHIM_CAT(MTK_SHIM_PLUS_A entries, _END)                                                     \
        }                                                                                                   \
        void boxminus(scalar* mtk_res_, const name& mtk_other_) const                                       \
        {                                                                                                   \
            MTK_SHIM_CAT(MTK_SHIM_MINUS_A entries, _END)                                                    \
        }                                                                                                   \
    };

#endif
