// TEST INFRASTRUCTURE (not product code).
#ifndef VLREF_EVENTS_H_
#define VLREF_EVENTS_H_
#include <stdint.h>
#include <string>
struct VlrefEvents {
    uint64_t finish_count = 0, first_finish_time = 0;
    uint64_t stop_count = 0, first_stop_time = 0;
    uint64_t fatal_count = 0, first_fatal_time = 0;
    int first_stop_line = 0, first_fatal_line = 0;
};
void vlref_set_events(VlrefEvents* p);
void vlref_set_capture(std::string* p);
#endif
