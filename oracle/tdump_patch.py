#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- patches the scratch copy of the reference that
oracle/Makefile's `tdump` target makes under oracle/_ref/tdump_src (nothing in
the repo or in /root/reference is touched):

  tracer/tracer.cpp  at XX_end (tracer.cpp:5161-5174) store ctx_T_BUF(0) of the
                     packet into the path tracer's red plane, with the index
                     arithmetic of tracer.cpp:5184-5190
  engine/engine.h    a public getter for that plane (rt_Scene::qr_tbuf)
"""
import sys

td = sys.argv[1]

p = td + "/tracer/tracer.cpp"
s = open(p).read()
anchor = "        movxx_ld(Rebx, Mebp, inf_FSAA)\n        movxx_ri(Resi, IM(RT_SIMD_QUADS*16))\n"
assert s.count(anchor) == 1, "XX_end anchor not found exactly once"
store = anchor + """
        /* QR_TDUMP: primary hit distance of the packet -> red plane */
        movxx_ld(Reax, Mebp, inf_FRM_Y)
        mulxx_ld(Reax, Mebp, inf_FRM_ROW)
        addxx_ld(Reax, Mebp, inf_FRM_X)
        shlxx_ri(Reax, IB(L+1))
        shlxx_rr(Reax, Rebx)
        movxx_ld(Redx, Mebp, inf_PTR_R)
        movpx_ld(Xmm0, Mecx, ctx_T_BUF(0))
        movpx_st(Xmm0, Iedx, DP(0))
"""
open(p, "w").write(s.replace(anchor, store))

p = td + "/engine/engine.h"
s = open(p).read()
anchor = "    rt_si32     get_x_row();\n"
assert s.count(anchor) == 1
open(p, "w").write(s.replace(anchor, anchor + "    rt_real    *qr_tbuf() { return ptr_r; }\n"))
