// The reference headers include <Mahi/Util/Timing/Time.hpp> (ModelParameters.hpp:3); everything lives in Util.hpp here.
#pragma once
#include <Mahi/Util.hpp>
