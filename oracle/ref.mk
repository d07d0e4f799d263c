# ref.mk — compiles the REFERENCE'S OWN sources, where they lie under /root/reference/src, against the
# stand-in headers of oracle/shims/ (ai.h, Eigen, cryptomatte.h) and the lens pack emitted in the
# generator's format, into oracle/_ref/libref.so.  Only runs where /root/reference exists (this
# container); the GPU box uses the prebuilt file.  Nothing is copied out of /root/reference.
# Flags as the reference's own build (CMakeLists.txt:17,28): C++17, -O3 -DNDEBUG.
PY ?= python
CXX ?= g++
REF ?= /root/reference
HERE := $(dir $(abspath $(lastword $(MAKEFILE_LIST))))
OUT := $(HERE)_ref
GENROOT := $(OUT)/gen
# include-path trick: the reference includes "../../Eigen/Eigen/Core", "../CryptomatteArnold/..." and
# "../../../polynomial-optics/..." relative to its own tree; those fall through to -I directories, where the
# leading ../ components are absorbed by dummy path segments
INC := -I$(HERE)shims -I$(HERE)shims/a/b -I$(HERE)shims/a -I$(GENROOT)/a/b/c -I$(REF)/src
CXXFLAGS ?= -std=c++17 -O3 -DNDEBUG -fPIC -ffp-contract=off -pthread -w
SRCS := $(REF)/src/lentil.cpp $(REF)/src/lentil_camera.cpp $(REF)/src/lentil_filter.cpp $(REF)/src/lentil_imager.cpp \
        $(REF)/src/lentil_operator.cpp $(REF)/src/lentil_loader.cpp
OBJS := $(OUT)/lentil.o $(OUT)/lentil_camera.o $(OUT)/lentil_filter.o $(OUT)/lentil_imager.o $(OUT)/lentil_operator.o $(OUT)/lentil_loader.o \
        $(OUT)/ref_harness.o
PACK := $(wildcard $(HERE)../pota_b200/lenses/*.json)

all: $(OUT)/libref.so

$(GENROOT)/.stamp: $(PACK) $(HERE)../pota_b200/lensgen/emit.py
	mkdir -p $(GENROOT)/a/b/c
	cd $(HERE).. && $(PY) -m pota_b200.lensgen.emit upstream $(GENROOT)
	touch $@

$(OUT)/%.o: $(REF)/src/%.cpp $(GENROOT)/.stamp $(HERE)shims/ai.h $(HERE)shims/Eigen/Eigen/Core
	$(CXX) $(CXXFLAGS) $(INC) -c $< -o $@

$(OUT)/ref_harness.o: $(HERE)ref_harness.cpp $(GENROOT)/.stamp $(HERE)shims/ai.h $(HERE)shims/Eigen/Eigen/Core $(HERE)../include/lentil_b200.h
	$(CXX) $(CXXFLAGS) $(INC) -c $< -o $@

$(OUT)/libref.so: $(OBJS)
	$(CXX) -shared -o $@ $(OBJS)

clean:
	rm -rf $(OUT)
.PHONY: all clean
