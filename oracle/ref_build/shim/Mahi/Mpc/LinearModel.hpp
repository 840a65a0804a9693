// The reference's DoublePendulumModel.hpp includes <Mahi/Mpc/LinearModel.hpp>; the file itself
// lives at old/LinearModel.hpp in the reference tree. Forward to it (include path set by Makefile).
#pragma once
#include <LinearModel.hpp>
