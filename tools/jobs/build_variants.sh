# build_variants.sh name1:"-DFLAG=.." name2:"..." -> build/variants/<name>.so (run HERE; the .so files travel to the GPU box)
mkdir -p build/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -prec-div=false -prec-sqrt=false -ftz=true -shared -Xcompiler -fPIC $flags \
      -o build/variants/$name.so pupperv3_mjx_b200/csrc/pupper_env.cu && echo built $name ) &
done
wait
