# oracle/ref_recipe.mk -- builds oracle/_ref/libref.so: the reference's own UKF sources, compiled UNMODIFIED from where
# they lie under /root/reference, against the stand-in dependency headers of oracle/ref_shim, behind the oracle's batch
# C interface (oracle/ref_capi.cpp).  TEST INFRASTRUCTURE: only tests/test_ref_pin.py loads the result, to check the
# CPU oracle against it.  Outputs go to oracle/_ref/ only (git-ignored, travels to the GPU box like other built .so).
# The reference's own build system (Rock CMake macros) is not used; nothing of the reference is copied into the repo --
# oracle/_ref/include/pose_estimation is a symlink to /root/reference/src so that its <pose_estimation/...> includes
# resolve.  Run from the repo root:  make -f oracle/ref_recipe.mk     (skipped with a message when /root/reference is absent)
REF ?= /root/reference
CXX := /usr/bin/g++
# the reference is C++03/11-era code built by Rock as plain x86-64 (no FMA contraction)
CXXFLAGS ?= -O2 -std=c++17 -ffp-contract=off -fopenmp -fPIC -w
OUT := oracle/_ref
INC := -I oracle/ref_shim -I $(OUT)/include -I include
SHIM := $(shell find oracle/ref_shim -type f)

ifeq ($(wildcard $(REF)/src/UnscentedKalmanFilter.hpp),)
all:
	@echo "oracle/_ref: $(REF) is not present here; keeping whatever oracle/_ref already holds"
else
all: $(OUT)/libref.so

$(OUT)/include/pose_estimation:
	@mkdir -p $(OUT)/include
	ln -sfn $(REF)/src $(OUT)/include/pose_estimation

$(OUT)/PoseUKF.o: $(REF)/src/pose_with_velocity/PoseUKF.cpp $(SHIM) oracle/ukf_oracle.hpp | $(OUT)/include/pose_estimation
	$(CXX) $(CXXFLAGS) $(INC) -c $< -o $@

$(OUT)/OrientationUKF.o: $(REF)/src/orientation_estimator/OrientationUKF.cpp $(SHIM) oracle/ukf_oracle.hpp | $(OUT)/include/pose_estimation
	$(CXX) $(CXXFLAGS) $(INC) -c $< -o $@

$(OUT)/ref_capi.o: oracle/ref_capi.cpp oracle/oracle_capi.cpp $(SHIM) oracle/ukf_oracle.hpp | $(OUT)/include/pose_estimation
	$(CXX) $(CXXFLAGS) $(INC) -I oracle -c $< -o $@

$(OUT)/libref.so: $(OUT)/PoseUKF.o $(OUT)/OrientationUKF.o $(OUT)/ref_capi.o
	$(CXX) $(CXXFLAGS) -shared -o $@ $^
	@sha256sum $(REF)/src/pose_with_velocity/PoseUKF.cpp $(REF)/src/orientation_estimator/OrientationUKF.cpp $(REF)/src/UnscentedKalmanFilter.hpp > $(OUT)/sources.sha256
endif
