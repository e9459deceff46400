#!/bin/bash
# developer helper: build csrc in a scratch copy (/tmp/b) and swap the finished .so into the tree atomically, so that a gpurun
# snapshot taken at any moment sees a consistent library.  usage: tools/install_tmp_build.sh [sync|build|install]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
case "$1" in
  sync)    rm -rf /tmp/b && mkdir -p /tmp/b/madnlp.jl_b200 && cp -r $ROOT/include /tmp/b/include && cp -rp $ROOT/madnlp.jl_b200/csrc /tmp/b/madnlp.jl_b200/csrc ;;
  build)   make -C /tmp/b/madnlp.jl_b200/csrc 2>&1 | grep -E "error|warning" -A6 | head -40; ls -la /tmp/b/madnlp.jl_b200/csrc/libb200kkt.so ;;
  install) cp -p /tmp/b/madnlp.jl_b200/csrc/*.cu /tmp/b/madnlp.jl_b200/csrc/*.cuh /tmp/b/madnlp.jl_b200/csrc/*.cpp /tmp/b/madnlp.jl_b200/csrc/*.hpp /tmp/b/madnlp.jl_b200/csrc/*.o $ROOT/madnlp.jl_b200/csrc/
           cp -p /tmp/b/include/*.h $ROOT/include/
           cp -p /tmp/b/madnlp.jl_b200/csrc/libb200kkt.so $ROOT/madnlp.jl_b200/csrc/libb200kkt.so.new && mv $ROOT/madnlp.jl_b200/csrc/libb200kkt.so.new $ROOT/madnlp.jl_b200/csrc/libb200kkt.so
           ls -la $ROOT/madnlp.jl_b200/csrc/libb200kkt.so ;;
esac
