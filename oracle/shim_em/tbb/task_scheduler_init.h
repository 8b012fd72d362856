#pragma once
namespace tbb { class task_scheduler_init { public: task_scheduler_init(int = 1) {} }; }
