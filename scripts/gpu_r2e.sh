for e in 16 8 4; do for w in 1 3; do timeout -s KILL 200 python benchmarks/dispatch_latency.py $w $e; done; done
