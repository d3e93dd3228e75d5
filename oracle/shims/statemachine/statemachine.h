/* ORACLE shim for ggplib's statemachine/statemachine.h: the abstract interface the reference's
 * search core calls (call-site census in SURVEY.md §1). The concrete implementation used by the
 * harness (oracle/ref_harness.cpp) forwards to oracle/propnet_ref.c. */
#pragma once

#include <string>

#include "basestate.h"
#include "jointmove.h"

namespace GGPLib {

class LegalState {
public:
    LegalState() : count(0), legals(nullptr) {}
    int getCount() const { return count; }
    int getLegal(int i) const { return legals[i]; }
    int count;
    const int* legals;
};

class StateMachineInterface {
public:
    virtual ~StateMachineInterface() {}
    virtual StateMachineInterface* dupe() const = 0;
    virtual BaseState* newBaseState() const = 0;
    virtual const BaseState* getInitialState() const = 0;
    virtual void updateBases(const BaseState* bs) = 0;
    virtual void reset() = 0;
    virtual LegalState* getLegalState(int role_index) = 0;
    virtual JointMove* getJointMove() = 0;
    virtual void nextState(const JointMove* move, BaseState* bs) = 0;
    virtual bool isTerminal() const = 0;
    virtual int getGoalValue(int role_index) = 0;
    virtual int getRoleCount() const = 0;
    virtual const char* legalToMove(int role_index, int choice) const = 0;
};

}  // namespace GGPLib
