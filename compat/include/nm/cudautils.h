// cudautils.h -- drop-in name of the reference header; the declarations live in nm_compat.hpp
#pragma once
#include "nm_compat.hpp"
