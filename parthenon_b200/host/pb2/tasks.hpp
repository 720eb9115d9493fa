// tasks.hpp — task lists / regions / collections on a single host thread.
//
// Same surface as the reference's src/tasks/tasks.hpp:158-330: TaskID with operator|,
// TaskList::AddTask(dependency, function, args...), TaskRegion (one list per partition),
// TaskCollection::Execute().  A task that returns TaskStatus::incomplete is polled again
// on the next sweep (tasks.cpp:129-139).  Tasks only ENQUEUE work on CUDA streams, so a
// whole stage is submitted in microseconds and the device never waits on the host.
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "types.hpp"

namespace parthenon {

class TaskID {
 public:
  TaskID() = default;
  explicit TaskID(int id) {
    if (id > 0) ids_.push_back(id);
  }
  TaskID operator|(const TaskID &o) const {
    TaskID out = *this;
    for (int i : o.ids_) out.ids_.push_back(i);
    return out;
  }
  const std::vector<int> &ids() const { return ids_; }
  bool empty() const { return ids_.empty(); }

 private:
  std::vector<int> ids_;
};

class TaskList {
 public:
  // tasks.hpp:197: callable + by-value arguments, like std::bind
  template <class F, class... Args>
  TaskID AddTask(const TaskID &dep, F &&func, Args &&...args) {
    auto bound = [f = std::forward<F>(func),
                  tup = std::make_tuple(std::forward<Args>(args)...)]() mutable -> TaskStatus {
      return std::apply(f, tup);
    };
    tasks_.push_back(Task{std::function<TaskStatus()>(std::move(bound)), dep.ids(), false});
    return TaskID(static_cast<int>(tasks_.size()));
  }
  // member-function form: AddTask(dep, &Class::method, object, args...)
  bool IsComplete() const {
    for (auto &t : tasks_)
      if (!t.done) return false;
    return true;
  }
  int Size() const { return static_cast<int>(tasks_.size()); }
  // run every task whose dependencies are complete; returns fail on the first failure
  TaskListStatus DoAvailable() {
    bool progressed = true;
    while (progressed) {
      progressed = false;
      for (auto &t : tasks_) {
        if (t.done) continue;
        bool ready = true;
        for (int d : t.deps) ready = ready && tasks_[d - 1].done;
        if (!ready) continue;
        const TaskStatus s = t.f();
        if (s == TaskStatus::fail) return TaskListStatus::stuck;
        if (s == TaskStatus::complete) {
          t.done = true;
          progressed = true;
        }
      }
    }
    return IsComplete() ? TaskListStatus::complete : TaskListStatus::running;
  }

 private:
  struct Task {
    std::function<TaskStatus()> f;
    std::vector<int> deps;
    bool done;
  };
  std::vector<Task> tasks_;
};

class TaskRegion {
 public:
  explicit TaskRegion(int n) : lists_(n) {}
  TaskList &operator[](int i) { return lists_[i]; }
  int size() const { return static_cast<int>(lists_.size()); }
  // tasks.cpp:119-150: sweep the lists until all are complete
  TaskListStatus Execute() {
    for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
      bool all = true;
      for (auto &tl : lists_) {
        if (tl.IsComplete()) continue;
        const TaskListStatus s = tl.DoAvailable();
        if (s == TaskListStatus::stuck) return TaskListStatus::stuck;
        all = all && (s == TaskListStatus::complete);
      }
      if (all) return TaskListStatus::complete;
    }
    return TaskListStatus::stuck;
  }

 private:
  static constexpr int kMaxSweeps = 100000000;
  std::vector<TaskList> lists_;
};

class TaskCollection {
 public:
  TaskRegion &AddRegion(int num_lists) {
    regions_.emplace_back(std::make_unique<TaskRegion>(num_lists));
    return *regions_.back();
  }
  TaskListStatus Execute() {
    for (auto &r : regions_) {
      const TaskListStatus s = r->Execute();
      if (s != TaskListStatus::complete) return s;
    }
    return TaskListStatus::complete;
  }

 private:
  std::vector<std::unique_ptr<TaskRegion>> regions_;
};

} // namespace parthenon
