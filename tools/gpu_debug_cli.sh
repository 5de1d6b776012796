cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out /tmp/dbg
for f in tests/golden/fuzz/f11 tests/golden/fuzz/f12 tests/golden/micro/tumor; do
  for env in "X=1" "SEEKSV_B200_CHUNK_LOG2=14" "SEEKSV_B200_CHUNK_LOG2=10"; do
      env $env seeksv_b200/bin/seeksv getclip -o /tmp/dbg/x $f.sort.bam > /tmp/dbg/out 2> /tmp/dbg/err; rc=$?
      echo "$f $env rc=$rc $(head -c 100 /tmp/dbg/err | tr '\n' ' ')"
  done
done
