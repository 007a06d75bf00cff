SPEC='aformat=sample_fmts=dbl,agate=threshold=0.037495:ratio=2.0:attack=5.00:release=200:range=0.1995:knee=3.0:detection=rms:makeup=1.0'
timeout 600 python -m pytest tests/test_gpu_filters.py -x -q 2>&1 | tail -3
for seg in 16384 36736 ; do echo seg $seg; JT_ENV_SEG=$seg timeout 300 python scripts/time_filter.py "$SPEC" 60 48000 f32; done
echo old; JT_NO_TILES=1 timeout 300 python scripts/time_filter.py "$SPEC" 60 48000 f32
