# chunk sizes of CHAINED streamed calls (tools/stream_bench.py): usage: bash tools/sweep_stream.sh [segments]
for cfg in "10,3" "12,3" "16,3" "16,4" "20,4" "24,3" "32,3" "64,2"; do
  c=${cfg%,*}; nb=${cfg#*,}
  echo "chunks=$c nbuf=$nb"
  PCSEG_HOST_SCHED=4,$c,$c,$nb PCSEG_HOST_CHUNKS=$c timeout 100 python tools/stream_bench.py 10 $1 2>/dev/null | grep streaming
done
