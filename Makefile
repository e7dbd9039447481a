# Build recipe for every native artefact in the repo (all in-tree, git-ignored):
#
#   minihost/libavs_minihost.so                      AviSynth+ C-API stand-in (tests/bench host)
#   oracle/libjinc_oracle.so                          plain-C restatement of the reference algorithm (checker)
#   oracle/_ref/libjincresize_ref.so                  UNMODIFIED reference, compiled where it lies (only when
#                                                     /root/reference is mounted; never copied into the repo)
#   avisynth-jincresize_b200/libjinc_b200.so          sm_100a kernels + the C ABI of include/jinc_b200.h
#   avisynth-jincresize_b200/libjincresize_b200.so    the AviSynth+ C plugin (drop-in for the reference .so)
#   avisynth-jincresize_b200/libvsjincresize_b200.so  the VapourSynth (API 4) front-end over the same C ABI
#   minihost/libvs_minihost.so                        VapourSynth core stand-in (tests)
#   avisynth-jincresize_b200/fma_peak                 FP32-FMA-pipe microbenchmark (roofline denominator)

CXX       ?= g++
CC        ?= gcc
NVCC      ?= /usr/local/cuda/bin/nvcc
CUDA_HOME ?= /usr/local/cuda
REF       ?= /root/reference/src

PKG   := avisynth-jincresize_b200
HOSTI := -Iminihost/include -Iminihost
# VapourSynth headers: the clean-room restatement by default; point VSINC at a real SDK for a production build
VSINC ?= minihost/include
CXXFLAGS_COMMON := -std=c++17 -fPIC -fvisibility=hidden -Wall -Wno-unused-function

ARCH  := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Iinclude -I$(PKG)/csrc

CUDA_SRCS := $(wildcard $(PKG)/csrc/*.cu)
CUDA_HOST_SRCS := $(wildcard $(PKG)/csrc/*.cpp)
CUDA_HDRS := $(wildcard $(PKG)/csrc/*.h $(PKG)/csrc/*.cuh) include/jinc_b200.h
CUDA_OBJS := $(CUDA_SRCS:.cu=.o) $(CUDA_HOST_SRCS:.cpp=.o)

.PHONY: all host cuda ref clean
all: host cuda ref
host: minihost/libavs_minihost.so minihost/libvs_minihost.so oracle/libjinc_oracle.so
cuda: $(PKG)/libjinc_b200.so $(PKG)/libjincresize_b200.so $(PKG)/libvsjincresize_b200.so $(PKG)/fma_peak

ifneq ($(wildcard $(REF)/JincResize.cpp),)
ref: oracle/_ref/libjincresize_ref.so
else
ref:
	@echo "reference sources not mounted at $(REF): keeping any prebuilt oracle/_ref"
endif

minihost/libavs_minihost.so: minihost/minihost.cpp minihost/minihost.h minihost/include/avisynth_c.h
	$(CXX) $(CXXFLAGS_COMMON) -O2 $(HOSTI) -shared -o $@ minihost/minihost.cpp -ldl -lpthread

minihost/libvs_minihost.so: minihost/vs_minihost.cpp minihost/include/VapourSynth4.h
	$(CXX) $(CXXFLAGS_COMMON) -O2 -Iminihost/include -shared -o $@ minihost/vs_minihost.cpp -ldl -lpthread

oracle/libjinc_oracle.so: oracle/jinc_oracle.c oracle/jinc_oracle.h
	$(CC) -std=c11 -O2 -fPIC -Wall -Werror=implicit-function-declaration -ffp-contract=off -shared -o $@ oracle/jinc_oracle.c -lm

# The reference's own CMake defaults to Release (-O3 -DNDEBUG), C++17, and gives the three SIMD files their ISA
# flags (CMakeLists.txt:3-7,57-61,65).  The same flags are used here; the sources are read from the mount.
REF_FLAGS := -std=gnu++17 -O3 -DNDEBUG -fPIC $(HOSTI) -I$(REF)
oracle/_ref/libjincresize_ref.so: oracle/ref_shim.cpp minihost/include/avisynth_c.h minihost/libavs_minihost.so
	mkdir -p oracle/_ref
	$(CXX) $(REF_FLAGS) -c $(REF)/JincResize.cpp -o oracle/_ref/JincResize.o
	$(CXX) $(REF_FLAGS) -msse4.1 -mfpmath=sse -c $(REF)/resize_plane_sse41.cpp -o oracle/_ref/resize_plane_sse41.o
	$(CXX) $(REF_FLAGS) -mavx2 -mfma -c $(REF)/resize_plane_avx2.cpp -o oracle/_ref/resize_plane_avx2.o
	$(CXX) $(REF_FLAGS) -mavx512f -mavx512bw -mavx512dq -mavx512vl -mfma -c $(REF)/resize_plane_avx512.cpp -o oracle/_ref/resize_plane_avx512.o
	$(CXX) $(REF_FLAGS) -c oracle/ref_shim.cpp -o oracle/_ref/ref_shim.o
	$(CXX) -shared -o $@ oracle/_ref/JincResize.o oracle/_ref/resize_plane_sse41.o oracle/_ref/resize_plane_avx2.o \
	    oracle/_ref/resize_plane_avx512.o oracle/_ref/ref_shim.o
	rm -f oracle/_ref/*.o

$(PKG)/csrc/%.o: $(PKG)/csrc/%.cu $(CUDA_HDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(PKG)/csrc/%.o: $(PKG)/csrc/%.cpp $(CUDA_HDRS)
	$(CXX) $(CXXFLAGS_COMMON) -O2 -ffp-contract=off -Iinclude -I$(PKG)/csrc -I$(CUDA_HOME)/include -c $< -o $@

$(PKG)/libjinc_b200.so: $(CUDA_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(CUDA_OBJS) -cudart shared -lpthread

$(PKG)/libjincresize_b200.so: $(PKG)/plugin/jincresize_plugin.cpp include/jinc_b200.h minihost/include/avisynth_c.h $(PKG)/libjinc_b200.so
	$(CXX) $(CXXFLAGS_COMMON) -O2 $(HOSTI) -Iinclude -shared -o $@ $(PKG)/plugin/jincresize_plugin.cpp \
	    -L$(PKG) -ljinc_b200 -Wl,-rpath,'$$ORIGIN' -lpthread

$(PKG)/libvsjincresize_b200.so: $(PKG)/vapoursynth/vsjincresize_plugin.cpp include/jinc_b200.h $(VSINC)/VapourSynth4.h $(PKG)/libjinc_b200.so
	$(CXX) $(CXXFLAGS_COMMON) -O2 -I$(VSINC) -Iinclude -shared -o $@ $(PKG)/vapoursynth/vsjincresize_plugin.cpp \
	    -L$(PKG) -ljinc_b200 -Wl,-rpath,'$$ORIGIN' -lpthread

$(PKG)/fma_peak: $(PKG)/tools/fma_peak.cu
	$(NVCC) $(ARCH) -std=c++17 -O3 -lineinfo -o $@ $<

clean:
	rm -f minihost/*.so oracle/*.so $(PKG)/*.so $(PKG)/csrc/*.o $(PKG)/fma_peak
	rm -rf oracle/_ref
