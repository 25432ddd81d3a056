# cmake/SEALConfig.cmake — makes `find_package(SEAL 4.1 REQUIRED)` (phanen/pplp CMakeLists.txt:29) resolve to pplp_b200.
#
#   cmake -S /path/to/pplp -B build -DSEAL_DIR=/path/to/pplp_b200/cmake
#
# Exports the imported target SEAL::seal (CMakeLists.txt:33-35 links it into pplp, client, server, tc, ts): the header-only
# SEAL-4.1 API subset in include/seal/seal.h over the C ABI of pplp_b200/libpplp_b200.so (hand-written sm_100a kernels).
# Build the library first: `python -m pplp_b200.build`.
get_filename_component(PPLP_B200_ROOT "${CMAKE_CURRENT_LIST_DIR}/.." ABSOLUTE)
set(PPLP_B200_LIBRARY "${PPLP_B200_ROOT}/pplp_b200/libpplp_b200.so")
if(NOT EXISTS "${PPLP_B200_LIBRARY}")
  set(SEAL_FOUND FALSE)
  set(SEAL_NOT_FOUND_MESSAGE "pplp_b200: ${PPLP_B200_LIBRARY} is missing - run `python -m pplp_b200.build` (nvcc, sm_100a) first")
  return()
endif()

find_package(ZLIB REQUIRED)

if(NOT TARGET SEAL::seal)
  add_library(SEAL::seal SHARED IMPORTED)
  set_target_properties(SEAL::seal PROPERTIES
    IMPORTED_LOCATION "${PPLP_B200_LIBRARY}"
    IMPORTED_NO_SONAME TRUE
    INTERFACE_INCLUDE_DIRECTORIES "${PPLP_B200_ROOT}/include"
    INTERFACE_LINK_LIBRARIES "ZLIB::ZLIB;${CMAKE_DL_LIBS}"
    INTERFACE_COMPILE_FEATURES cxx_std_17)
  # src/demo.cc includes include/bloomfilter.h (which uses uint8_t) before any header that declares it; GCC >= 13 no
  # longer leaks <cstdint> through <sstream>.  A toolchain matter of the reference, independent of the SEAL provider.
  if(CMAKE_CXX_COMPILER_ID MATCHES "GNU|Clang")
    set_property(TARGET SEAL::seal APPEND PROPERTY INTERFACE_COMPILE_OPTIONS "SHELL:-include cstdint")
  endif()
endif()

set(SEAL_FOUND TRUE)
set(SEAL_VERSION 4.1.1)
set(SEAL_VERSION_MAJOR 4)
set(SEAL_VERSION_MINOR 1)
set(SEAL_VERSION_PATCH 1)
# what upstream SEALConfig.cmake reports about its build; the reference reads none of them
set(SEAL_BUILD_TYPE Release)
set(SEAL_USE_CXX17 ON)
set(SEAL_USE_ZLIB ON)
set(SEAL_USE_ZSTD ON)      # zstd streams are read and written through the runtime library (dlopen)
set(SEAL_USE_INTEL_HEXL OFF)
