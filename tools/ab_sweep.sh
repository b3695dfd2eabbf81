for pb in 1 2 4; do python tools/var_bench.py lib_pb$pb.so:::1024 lib_pb$pb.so:::2368; done
