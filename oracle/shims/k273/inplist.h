/* ORACLE shim for k273/inplist.h: the minimal intrusive list the (stubbed) LRU cache declares
 * (src/cpp/scheduler.h:49, scheduler.cpp:58-62,73-85). */
#pragma once
#include <utility>

namespace K273 {
template <typename T>
class InplaceList {
public:
    struct Node {
        Node* prev = nullptr;
        Node* next = nullptr;
        T value;
        T& get() { return value; }
    };
    InplaceList() : head_(nullptr), tail_(nullptr), count_(0) {}
    ~InplaceList() { while (head_) { Node* n = head_; head_ = n->next; delete n; } }
    template <typename... Args>
    Node* emplaceBack(Args&&... args) {
        Node* n = new Node();
        n->value = T(std::forward<Args>(args)...);
        n->prev = tail_;
        if (tail_) tail_->next = n; else head_ = n;
        tail_ = n;
        count_++;
        return n;
    }
    void remove(Node* n) {
        if (n->prev) n->prev->next = n->next; else head_ = n->next;
        if (n->next) n->next->prev = n->prev; else tail_ = n->prev;
        n->prev = n->next = nullptr;
        count_--;
    }
    void pushFront(Node* n) {
        n->prev = nullptr; n->next = head_;
        if (head_) head_->prev = n; else tail_ = n;
        head_ = n;
        count_++;
    }
    Node* head() { return head_; }
    bool empty() const { return count_ == 0; }
    int size() const { return count_; }
private:
    Node* head_;
    Node* tail_;
    int count_;
};
}  // namespace K273
