// include/compat/test.h — source-compatible test registry of the reference (include/test.h:7-22): TEST(name)
// defines and registers a test, SKIP(name) only defines it, test(wildcard) runs the registered ones whose name
// matches the regular expression.
#pragma once
#ifndef RMD_COMPAT_TEST_H
#define RMD_COMPAT_TEST_H

#include <functional>
#include <string>
#include <utility>
#include <vector>

typedef std::vector<std::pair<std::string, std::function<void()>>> FuncVector;
extern FuncVector registered_funcs;

#define TEST(func_name)                                                                   \
    void func_name();                                                                     \
    struct func_name##_registrar {                                                        \
        func_name##_registrar() { registered_funcs.push_back({#func_name, func_name}); }  \
    } func_name##_instance;                                                               \
    void func_name()

#define SKIP(func_name) void func_name()

void test(std::string wildcard = ".*");

#endif
