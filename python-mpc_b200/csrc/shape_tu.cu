// One translation unit per compiled (shape, dtype):  nvcc -c shape_tu.cu -DMPCB_NX=5 -DMPCB_NU=1 -DMPCB_SLACK=1 -DMPCB_TU_F64=1
// registers its ShapeOps table with mpc_b200.cu when the library is loaded (see __graft_entry__.SHAPES).
#include "shape_ops.cuh"

#if !defined(MPCB_NX) || !defined(MPCB_NU) || !defined(MPCB_SLACK) || !defined(MPCB_TU_F64)
#error "compile with -DMPCB_NX= -DMPCB_NU= -DMPCB_SLACK= -DMPCB_TU_F64="
#endif
#if MPCB_TU_F64
typedef double tu_real;
#else
typedef float tu_real;
#endif

namespace {
struct Registrar {
    Registrar() {
        static const ShapeOps ops = make_shape_ops<tu_real, Lay<MPCB_NX, MPCB_NU, (MPCB_SLACK != 0)>>();
        mpcb_rt::register_shape_ops(&ops);
    }
} registrar;      // runs when libmpc_b200.so is loaded
}  // namespace
