// Filters of the path whose kernels are not built yet fail loudly (no CPU fallback).
#include "jt_internal.h"
Sig jt_anlmdn(jt_ctx *, const Sig &, double, double, double, double) { JT_THROW(JT_ERR_UNSUPPORTED, "anlmdn kernel not built yet"); }
Sig jt_afftdn(jt_ctx *, const Sig &, const AfftdnParams &) { JT_THROW(JT_ERR_UNSUPPORTED, "afftdn kernel not built yet"); }
Sig jt_agate(jt_ctx *, const Sig &, const GateParams &) { JT_THROW(JT_ERR_UNSUPPORTED, "agate kernel not built yet"); }
Sig jt_acompressor(jt_ctx *, const Sig &, const CompParams &) { JT_THROW(JT_ERR_UNSUPPORTED, "acompressor kernel not built yet"); }
Sig jt_deesser(jt_ctx *, const Sig &, double, double, double) { JT_THROW(JT_ERR_UNSUPPORTED, "deesser kernel not built yet"); }
Sig jt_alimiter(jt_ctx *, const Sig &, const LimiterParams &) { JT_THROW(JT_ERR_UNSUPPORTED, "alimiter kernel not built yet"); }
Sig jt_adeclick(jt_ctx *, const Sig &, double, double, double, double, double, int) { JT_THROW(JT_ERR_UNSUPPORTED, "adeclick kernel not built yet"); }
