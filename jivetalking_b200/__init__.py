"""jivetalking-b200: B200-native replacement for jivetalking's four-pass DSP chain.
The product is jivetalking_b200/libjtdsp.so (C ABI in include/jtdsp.h); this package only
holds the ctypes binding used by tests and bench.py, and the synthetic input generator."""
