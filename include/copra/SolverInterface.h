// Drop-in include name of the reference (copra/SolverInterface.h); the implementation lives in copra_b200_facade.hpp.
#pragma once
#include "copra_b200_facade.hpp"
