"""orbx — host-side Python harness over the C-ABI library (include/orbx.h)."""
