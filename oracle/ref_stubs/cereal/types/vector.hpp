// stub: the oracle build never serialises (see cereal/archives/binary.hpp)
#pragma once
#include <vector>
