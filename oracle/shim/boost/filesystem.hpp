#pragma once
#include <filesystem>
namespace boost { namespace filesystem { using namespace std::filesystem; } }
