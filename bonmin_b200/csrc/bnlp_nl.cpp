#include "bnlp_internal.h"
namespace bnlp { bool parse_nl(const char *, size_t, HostModel &, std::string &err) { err = "not implemented"; return false; } }
