// see ../filtering_stream.hpp (the stand-in lives there)
#pragma once
#include "../filtering_stream.hpp"
