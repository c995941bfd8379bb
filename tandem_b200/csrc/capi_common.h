// Error plumbing of the C ABI: exceptions never cross the boundary.
#pragma once
#include <exception>
#include <string>

namespace tdm {
void set_last_error(const std::string& s);
}

#define TDM_API_BEGIN try {
#define TDM_API_END                                   \
  }                                                   \
  catch (const std::exception& e) {                   \
    ::tdm::set_last_error(e.what());                  \
    return TDM_ERR;                                   \
  }                                                   \
  catch (...) {                                       \
    ::tdm::set_last_error("unknown exception");       \
    return TDM_ERR;                                   \
  }
