# Package configuration of the B200-native NiftyMatch drop-in.  Installed next to the headers
# (<prefix>/include/nm), so consumers keep pointing NiftyMatch_DIR there and keep calling
# FIND_PACKAGE(NiftyMatch CONFIG REQUIRED).  Variables, as before:
#   NiftyMatch_INCLUDE_DIR, NiftyMatch_LIBS,
#   NiftyMatch_gpuutils_LIB, NiftyMatch_kernels_LIB, NiftyMatch_sift_LIB, NiftyMatch_PATH_SUFFIX
# New: NiftyMatch_core_LIB = the C-ABI library the three wrappers call (already part of NiftyMatch_LIBS).
set(NiftyMatch_PATH_SUFFIX nm)
get_filename_component(_nm_prefix "${CMAKE_CURRENT_LIST_DIR}/../.." ABSOLUTE)

find_path(NiftyMatch_INCLUDE_DIR NAMES macros.h
          PATHS "${_nm_prefix}/include" PATH_SUFFIXES ${NiftyMatch_PATH_SUFFIX} NO_DEFAULT_PATH)
foreach(_nm_lib gpuutils kernels sift)
    find_library(NiftyMatch_${_nm_lib}_LIB NAMES ${_nm_lib}
                 PATHS "${_nm_prefix}/lib" PATH_SUFFIXES ${NiftyMatch_PATH_SUFFIX} NO_DEFAULT_PATH)
endforeach()
find_library(NiftyMatch_core_LIB NAMES nm_b200
             PATHS "${_nm_prefix}/lib" PATH_SUFFIXES ${NiftyMatch_PATH_SUFFIX} NO_DEFAULT_PATH)

# order matters for static linking: sift -> kernels -> gpuutils -> core
set(NiftyMatch_LIBS ${NiftyMatch_gpuutils_LIB} ${NiftyMatch_kernels_LIB} ${NiftyMatch_sift_LIB}
                    ${NiftyMatch_kernels_LIB} ${NiftyMatch_gpuutils_LIB} ${NiftyMatch_core_LIB})

include(FindPackageHandleStandardArgs)
find_package_handle_standard_args(NiftyMatch DEFAULT_MSG NiftyMatch_LIBS NiftyMatch_INCLUDE_DIR
                                  NiftyMatch_gpuutils_LIB NiftyMatch_kernels_LIB NiftyMatch_sift_LIB
                                  NiftyMatch_core_LIB)
unset(_nm_prefix)
