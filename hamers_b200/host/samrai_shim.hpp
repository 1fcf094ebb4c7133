/*
 * samrai_shim.hpp -- the few SAMRAI types the hot path touches, for builds WITHOUT SAMRAI.
 *
 * HAMeRS reaches the convective-flux path through SAMRAI containers (hier::Patch, pdat::CellData / SideData,
 * geom::CartesianPatchGeometry, tbox::Database).  SAMRAI (v4.1.0, circleci/install-SAMRAI.sh:4) is not available in
 * this image, so this header restates -- from SAMRAI's documented behaviour, nothing copied -- exactly the part of
 * their interface that ConvectiveFluxReconstructor / RungeKuttaPatchStrategy use, with the SAME names and memory
 * layout (column-major, x fastest, one contiguous array per depth component; side data: one array per normal
 * direction and component whose extent is N+1 along the normal).  With SAMRAI present, compile
 * with -DHAMERS_B200_WITH_SAMRAI and the real headers are used instead; the classes in this directory are written
 * against the common subset.
 */
#ifndef HAMERS_B200_SAMRAI_SHIM_HPP
#define HAMERS_B200_SAMRAI_SHIM_HPP

#ifdef HAMERS_B200_WITH_SAMRAI
#include "SAMRAI/geom/CartesianGridGeometry.h"
#include "SAMRAI/geom/CartesianPatchGeometry.h"
#include "SAMRAI/hier/IntVector.h"
#include "SAMRAI/hier/Patch.h"
#include "SAMRAI/hier/VariableContext.h"
#include "SAMRAI/pdat/CellData.h"
#include "SAMRAI/pdat/CellVariable.h"
#include "SAMRAI/pdat/SideData.h"
#include "SAMRAI/pdat/SideVariable.h"
#include "SAMRAI/tbox/Database.h"
#include "SAMRAI/tbox/Dimension.h"
#include "SAMRAI/tbox/Utilities.h"
#else

#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

/* The reference's only error convention: TBOX_ERROR prints and aborts (MPI_Abort).  Without SAMRAI it throws, so a
 * host program (and the tests) can observe it. */
#define TBOX_ERROR(msg)                                   \
    do {                                                  \
        std::ostringstream tbox_os_;                      \
        tbox_os_ << msg;                                  \
        throw std::runtime_error(tbox_os_.str());         \
    } while (0)
#define TBOX_ASSERT(x)                                                     \
    do {                                                                   \
        if (!(x)) TBOX_ERROR("Failed assertion: " #x << std::endl);        \
    } while (0)
#define NULL_USE(x) (void)(x)

namespace SAMRAI {

namespace tbox {

class Dimension {
public:
    explicit Dimension(unsigned short d) : d_dim(d) {}
    unsigned short getValue() const { return d_dim; }
    bool operator==(const Dimension& o) const { return d_dim == o.d_dim; }
    bool operator>(const Dimension& o) const { return d_dim > o.d_dim; }

private:
    unsigned short d_dim;
};

/* key/value store standing in for tbox::Database (input and restart databases) */
class Database {
public:
    explicit Database(const std::string& name = "") : d_name(name) {}
    bool keyExists(const std::string& k) const { return d_int.count(k) || d_dbl.count(k) || d_str.count(k) || d_vec.count(k); }
    int getIntegerWithDefault(const std::string& k, int dflt) const
    {
        auto it = d_int.find(k);
        return it == d_int.end() ? dflt : it->second;
    }
    int getInteger(const std::string& k) const
    {
        auto it = d_int.find(k);
        if (it == d_int.end()) TBOX_ERROR("Database '" << d_name << "': key '" << k << "' not found" << std::endl);
        return it->second;
    }
    void putInteger(const std::string& k, int v) { d_int[k] = v; }
    double getDoubleWithDefault(const std::string& k, double dflt) const
    {
        auto it = d_dbl.find(k);
        return it == d_dbl.end() ? dflt : it->second;
    }
    void putDouble(const std::string& k, double v) { d_dbl[k] = v; }
    std::vector<double> getDoubleVector(const std::string& k) const
    {
        auto it = d_vec.find(k);
        if (it == d_vec.end()) TBOX_ERROR("Database '" << d_name << "': key '" << k << "' not found" << std::endl);
        return it->second;
    }
    void putDoubleVector(const std::string& k, const std::vector<double>& v) { d_vec[k] = v; }
    std::string getStringWithDefault(const std::string& k, const std::string& dflt) const
    {
        auto it = d_str.find(k);
        return it == d_str.end() ? dflt : it->second;
    }
    void putString(const std::string& k, const std::string& v) { d_str[k] = v; }

private:
    std::string d_name;
    std::map<std::string, int> d_int;
    std::map<std::string, double> d_dbl;
    std::map<std::string, std::string> d_str;
    std::map<std::string, std::vector<double>> d_vec;
};

}  // namespace tbox

namespace hier {

class IntVector {
public:
    IntVector(const tbox::Dimension& dim, int v = 0) : d_dim(dim) { for (int i = 0; i < 3; i++) d_v[i] = i < dim.getValue() ? v : 0; }
    static IntVector getZero(const tbox::Dimension& dim) { return IntVector(dim, 0); }
    static IntVector getOne(const tbox::Dimension& dim) { return IntVector(dim, 1); }
    int& operator[](int i) { return d_v[i]; }
    const int& operator[](int i) const { return d_v[i]; }
    IntVector operator*(int s) const
    {
        IntVector r(*this);
        for (int i = 0; i < d_dim.getValue(); i++) r.d_v[i] *= s;
        return r;
    }
    bool operator==(const IntVector& o) const { return d_v[0] == o.d_v[0] && d_v[1] == o.d_v[1] && d_v[2] == o.d_v[2]; }
    const tbox::Dimension& getDim() const { return d_dim; }

private:
    tbox::Dimension d_dim;
    int d_v[3];
};

/* cell-centred index box [lower, upper], both inclusive */
class Box {
public:
    Box(const IntVector& lo, const IntVector& hi) : d_lo(lo), d_hi(hi) {}
    const IntVector& lower() const { return d_lo; }
    const IntVector& upper() const { return d_hi; }
    IntVector numberCells() const
    {
        IntVector n(d_lo);
        for (int i = 0; i < d_lo.getDim().getValue(); i++) n[i] = d_hi[i] - d_lo[i] + 1;
        return n;
    }
    void grow(const IntVector& g)
    {
        for (int i = 0; i < d_lo.getDim().getValue(); i++) {
            d_lo[i] -= g[i];
            d_hi[i] += g[i];
        }
    }
    const tbox::Dimension& getDim() const { return d_lo.getDim(); }

private:
    IntVector d_lo, d_hi;
};

class PatchData {
public:
    virtual ~PatchData() {}
};

class Variable {
public:
    Variable(const tbox::Dimension& dim, const std::string& name) : d_dim(dim), d_name(name) {}
    virtual ~Variable() {}
    const std::string& getName() const { return d_name; }
    const tbox::Dimension& getDim() const { return d_dim; }

private:
    tbox::Dimension d_dim;
    std::string d_name;
};

class VariableContext {
public:
    explicit VariableContext(const std::string& name) : d_name(name) {}
    const std::string& getName() const { return d_name; }

private:
    std::string d_name;
};

class PatchGeometry {
public:
    virtual ~PatchGeometry() {}
};

class Patch {
public:
    explicit Patch(const Box& box) : d_box(box) {}
    const Box& getBox() const { return d_box; }
    const tbox::Dimension& getDim() const { return d_box.getDim(); }
    void setPatchGeometry(const std::shared_ptr<PatchGeometry>& g) { d_geom = g; }
    std::shared_ptr<PatchGeometry> getPatchGeometry() const { return d_geom; }
    void setPatchData(const std::shared_ptr<Variable>& v, const std::shared_ptr<VariableContext>& c,
                      const std::shared_ptr<PatchData>& d)
    {
        d_data[std::make_pair(v.get(), c.get())] = d;
    }
    std::shared_ptr<PatchData> getPatchData(const std::shared_ptr<Variable>& v, const std::shared_ptr<VariableContext>& c) const
    {
        auto it = d_data.find(std::make_pair(v.get(), c.get()));
        if (it == d_data.end())
            TBOX_ERROR("Patch::getPatchData: variable '" << v->getName() << "' is not allocated in context '" << c->getName()
                                                          << "'" << std::endl);
        return it->second;
    }

private:
    Box d_box;
    std::shared_ptr<PatchGeometry> d_geom;
    std::map<std::pair<const Variable*, const VariableContext*>, std::shared_ptr<PatchData>> d_data;
};

}  // namespace hier

namespace geom {

class CartesianPatchGeometry : public hier::PatchGeometry {
public:
    CartesianPatchGeometry(const double* dx, const double* x_lo, int dim)
    {
        for (int i = 0; i < 3; i++) {
            d_dx[i] = i < dim ? dx[i] : 0.0;
            d_xlo[i] = (i < dim && x_lo) ? x_lo[i] : 0.0;
        }
    }
    const double* getDx() const { return d_dx; }
    const double* getXLower() const { return d_xlo; }

private:
    double d_dx[3], d_xlo[3];
};

class CartesianGridGeometry {
public:
    explicit CartesianGridGeometry(const tbox::Dimension& dim) : d_dim(dim) {}
    const tbox::Dimension& getDim() const { return d_dim; }

private:
    tbox::Dimension d_dim;
};

}  // namespace geom

namespace pdat {

/* CellData(box, depth, ghosts): per depth component one contiguous column-major array over the ghost box */
template <class T>
class CellData : public hier::PatchData {
public:
    CellData(const hier::Box& box, int depth, const hier::IntVector& ghosts) : d_box(box), d_ghost_box(box), d_depth(depth), d_ghosts(ghosts)
    {
        d_ghost_box.grow(ghosts);
        const hier::IntVector n = d_ghost_box.numberCells();
        d_size = 1;
        for (int i = 0; i < box.getDim().getValue(); i++) d_size *= (size_t)n[i];
        d_array.assign(d_size * depth, T(0));
    }
    T* getPointer(int d = 0) { return d_array.data() + d_size * d; }
    const T* getPointer(int d = 0) const { return d_array.data() + d_size * d; }
    int getDepth() const { return d_depth; }
    const hier::IntVector& getGhostCellWidth() const { return d_ghosts; }
    const hier::Box& getBox() const { return d_box; }
    const hier::Box& getGhostBox() const { return d_ghost_box; }
    void fillAll(const T& v) { d_array.assign(d_array.size(), v); }

private:
    hier::Box d_box, d_ghost_box;
    int d_depth;
    hier::IntVector d_ghosts;
    size_t d_size;
    std::vector<T> d_array;
};

/* SideData(box, depth, ghosts): per normal direction n and component d one array, extent N+1+2g along n */
template <class T>
class SideData : public hier::PatchData {
public:
    SideData(const hier::Box& box, int depth, const hier::IntVector& ghosts) : d_box(box), d_depth(depth), d_ghosts(ghosts)
    {
        const int dim = box.getDim().getValue();
        hier::Box gb(box);
        gb.grow(ghosts);
        const hier::IntVector n = gb.numberCells();
        for (int nd = 0; nd < dim; nd++) {
            size_t sz = 1;
            for (int i = 0; i < dim; i++) sz *= (size_t)(n[i] + (i == nd ? 1 : 0));
            d_size[nd] = sz;
            d_array[nd].assign(sz * depth, T(0));
        }
    }
    T* getPointer(int side_normal, int d = 0) { return d_array[side_normal].data() + d_size[side_normal] * d; }
    const T* getPointer(int side_normal, int d = 0) const { return d_array[side_normal].data() + d_size[side_normal] * d; }
    int getDepth() const { return d_depth; }
    const hier::IntVector& getGhostCellWidth() const { return d_ghosts; }
    const hier::Box& getBox() const { return d_box; }

private:
    hier::Box d_box;
    int d_depth;
    hier::IntVector d_ghosts;
    size_t d_size[3];
    std::vector<T> d_array[3];
};

template <class T>
class CellVariable : public hier::Variable {
public:
    CellVariable(const tbox::Dimension& dim, const std::string& name, int depth = 1) : hier::Variable(dim, name), d_depth(depth) {}
    int getDepth() const { return d_depth; }

private:
    int d_depth;
};

template <class T>
class SideVariable : public hier::Variable {
public:
    SideVariable(const tbox::Dimension& dim, const std::string& name, int depth = 1) : hier::Variable(dim, name), d_depth(depth) {}
    int getDepth() const { return d_depth; }

private:
    int d_depth;
};

}  // namespace pdat
}  // namespace SAMRAI

#endif /* HAMERS_B200_WITH_SAMRAI */

#include <memory>
#define HAMERS_SHARED_PTR std::shared_ptr
#define HAMERS_SHARED_PTR_CAST std::dynamic_pointer_cast
#define HAMERS_DYNAMIC_POINTER_CAST std::dynamic_pointer_cast

#endif /* HAMERS_B200_SAMRAI_SHIM_HPP */
