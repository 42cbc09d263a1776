// tbk_kdtree.cuh -- the neighbour choice of the mesh IDW fill, restated so that ties come out as the reference's.
//
// photometry/backgrounds.py:200-205 -> photutils 1.3.0 Background2D._interpolate_meshes ->
// ShepardIDWInterpolator(yx_good, values)  [scipy.spatial.cKDTree(yx_good, leafsize=10)]
//   .__call__(all mesh positions, n_neighbors=10, power=1, reg=0)  [kdtree.query(pos, k=10, eps=0)].
// The good meshes sit on an integer lattice, so many candidates are equidistant and WHICH of them make the
// ten depends on the kd-tree itself: the order of the points inside the leaves (left behind by the median
// selection + partition of the build) and the order in which the best-first search reaches the leaves.
// This header restates both, for lattice points, in integer arithmetic:
//   * build: balanced_tree=True, compact_nodes=True (the cKDTree defaults): every node recomputes its bounding box,
//     splits the dimension of largest extent at the coordinate of the median element (compared by that coordinate
//     alone), then partitions "< split | >= split" with the two-pointer sweep, sliding the
//     split when one side would be empty.  The median selection is the introspective selection of libstdc++
//     (median-of-three pivot moved to the front, unguarded Hoare partition, insertion sort below 4 elements, heap
//     selection when the depth limit 2 lg n runs out) -- scipy's wheels call std::nth_element, and the arrangement
//     it leaves behind decides the leaf order.
//   * query: scipy's query_single_point -- best-first traversal with a binary min-heap of (cell distance, cell)
//     whose sift rules are reproduced literally, near child first, a far child is queued only while its distance
//     does not exceed the current k-th distance, candidates replace the current worst only when strictly closer.
// Points are identified by their mesh id g = iy * nx + ix; the good-point order of the reference is increasing g,
// so "point index a < b" is "mesh id a < b".  All distances are squared integer distances (exact).
//
// tests/test_kdtree.py checks the tree (index permutation, node structure) and the neighbour sets against the
// real scipy.spatial.cKDTree on the CPU through tbk_debug_idw_neighbors.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define KDT_HD __host__ __device__ __forceinline__
#else
#define KDT_HD inline
#endif

#define KDT_LEAFSIZE 10      // photutils ShepardIDWInterpolator(leafsize=10)
#define KDT_K 10             // Background2D._interpolate_meshes(n_neighbors=10)
#define KDT_QCAP 96          // capacity of the cell queue of one search
#define KDT_INF 0x7fffffff

struct KdtNode {
	uint16_t a, b;   // leaf: points idx[a .. b);  inner node: children a (coordinate < split) and b (>= split)
	int16_t split;
	int16_t dim;     // -1 = leaf, 0 = mesh row (iy), 1 = mesh column (ix)
};

struct KdtTree {
	uint32_t* idx;       // [npts] permutation of the good meshes, each packed as (iy << 22) | (ix << 12) | mesh id
	KdtNode* nodes;      // [<= 2 npts]
	int npts, nnodes, nx;
	int mins[2], maxes[2];
	int overflow;        // node array or build stack exhausted (never for npts <= 4096 with the sizes used here)
};

KDT_HD uint32_t kdt_pack(int g, int nx) { return ((uint32_t)(g / nx) << 22) | ((uint32_t)(g % nx) << 12) | (uint32_t)g; }
KDT_HD int kdt_id(uint32_t w) { return (int)(w & 0xFFFu); }
// order of the median selection: the coordinate along d alone (equal coordinates are equivalent -- where they end up
// is decided by the selection algorithm, which is why it is restated step by step below)
KDT_HD int kdt_key(uint32_t w, int d) { return d == 0 ? (int)(w >> 22) : (int)((w >> 12) & 0x3FFu); }

// ---- selection: after the call p[nth] is the element of rank nth, smaller keys before it, larger after it -------
KDT_HD void kdt_swap(uint32_t& x, uint32_t& y) { const uint32_t t = x; x = y; y = t; }

KDT_HD void kdt_sift_down(uint32_t* p, int hole, int len, uint32_t value, int d)
{
	// max-heap on the key: move the hole down along the larger child, then bubble the value up from there
	const int top = hole;
	int child = hole;
	while (child < (len - 1) / 2) {
		child = 2 * (child + 1);
		if (kdt_key(p[child], d) < kdt_key(p[child - 1], d)) --child;
		p[hole] = p[child];
		hole = child;
	}
	if ((len & 1) == 0 && child == (len - 2) / 2) {
		child = 2 * (child + 1);
		p[hole] = p[child - 1];
		hole = child - 1;
	}
	int parent = (hole - 1) / 2;
	while (hole > top && kdt_key(p[parent], d) < kdt_key(value, d)) {
		p[hole] = p[parent];
		hole = parent;
		parent = (hole - 1) / 2;
	}
	p[hole] = value;
}

KDT_HD void kdt_heap_select(uint32_t* p, int middle, int last, int d)
{
	// heap of the `middle` smallest at the front
	if (middle >= 2) {
		for (int parent = (middle - 2) / 2; ; --parent) {
			kdt_sift_down(p, parent, middle, p[parent], d);
			if (parent == 0) break;
		}
	}
	for (int i = middle; i < last; ++i) {
		if (kdt_key(p[i], d) < kdt_key(p[0], d)) {
			const uint32_t v = p[i];
			p[i] = p[0];
			kdt_sift_down(p, 0, middle, v, d);
		}
	}
}

KDT_HD void kdt_nth_element(uint32_t* p, int nth, int n, int d)
{
	if (n == 0 || nth == n) return;
	int first = 0, last = n;
	int depth = 0;
	for (int m = n; m > 1; m >>= 1) ++depth;   // floor(lg n)
	depth *= 2;
	while (last - first > 3) {
		if (depth == 0) {
			kdt_heap_select(p + first, nth + 1 - first, last - first, d);
			kdt_swap(p[first], p[nth]);
			return;
		}
		--depth;
		// median of (first+1, mid, last-1) goes to the front as pivot
		const int mid = first + (last - first) / 2;
		{
			const int a = first + 1, b = mid, c = last - 1;
			const int ka = kdt_key(p[a], d), kb = kdt_key(p[b], d), kc = kdt_key(p[c], d);
			int m3;
			if (ka < kb) m3 = (kb < kc) ? b : ((ka < kc) ? c : a);
			else m3 = (ka < kc) ? a : ((kb < kc) ? c : b);
			kdt_swap(p[first], p[m3]);
		}
		const int kp = kdt_key(p[first], d);
		int lo = first + 1, hi = last;
		for (;;) {
			while (kdt_key(p[lo], d) < kp) ++lo;
			--hi;
			while (kp < kdt_key(p[hi], d)) --hi;
			if (!(lo < hi)) break;
			kdt_swap(p[lo], p[hi]);
			++lo;
		}
		if (lo <= nth) first = lo; else last = lo;
	}
	// insertion sort of the last <= 3 elements
	for (int i = first + 1; i < last; ++i) {
		const uint32_t v = p[i];
		const int kv = kdt_key(v, d);
		int j = i;
		while (j > first && kv < kdt_key(p[j - 1], d)) { p[j] = p[j - 1]; --j; }
		p[j] = v;
	}
}

// ---- build --------------------------------------------------------------------------------------------------
// One node: points idx[s .. e).  Returns false for a leaf; otherwise rearranges the points and returns the split
// dimension / value and the position p of the first point of the ">= split" side.
KDT_HD bool kdt_split(uint32_t* idx, int s, int e, int& p, int& dim, int& split)
{
	if (e - s <= KDT_LEAFSIZE) return false;
	int mn[2] = {KDT_INF, KDT_INF}, mx[2] = {-1, -1};
	for (int i = s; i < e; ++i) {
		const uint32_t w = idx[i];
		const int c0 = kdt_key(w, 0), c1 = kdt_key(w, 1);
		if (c0 < mn[0]) mn[0] = c0;
		if (c0 > mx[0]) mx[0] = c0;
		if (c1 < mn[1]) mn[1] = c1;
		if (c1 > mx[1]) mx[1] = c1;
	}
	int d = 0, size = 0;
	for (int i = 0; i < 2; ++i) if (mx[i] - mn[i] > size) { d = i; size = mx[i] - mn[i]; }
	if (mx[d] == mn[d]) return false;   // all points identical (cannot happen on a lattice)
	const int n = e - s;
	kdt_nth_element(idx + s, n / 2, n, d);
	split = kdt_key(idx[s + n / 2], d);
	p = s;
	int q = e - 1;
	while (p <= q) {
		if (kdt_key(idx[p], d) < split) ++p;
		else if (kdt_key(idx[q], d) >= split) --q;
		else { kdt_swap(idx[p], idx[q]); ++p; --q; }
	}
	if (p == s) {
		// no point below the split: the smallest coordinate becomes the split and goes left alone
		int j = s;
		split = kdt_key(idx[j], d);
		for (int i = s + 1; i < e; ++i) {
			const int c = kdt_key(idx[i], d);
			if (c < split) { j = i; split = c; }
		}
		kdt_swap(idx[s], idx[j]);
		p = s + 1;
	} else if (p == e) {
		int j = e - 1;
		split = kdt_key(idx[j], d);
		for (int i = s; i < e - 1; ++i) {
			const int c = kdt_key(idx[i], d);
			if (c > split) { j = i; split = c; }
		}
		kdt_swap(idx[e - 1], idx[j]);
		p = e - 1;
	}
	dim = d;
	return true;
}

// bounding box of all points (the tree's mins / maxes, which seed the cell distances of a search)
KDT_HD void kdt_root_box(KdtTree& t)
{
	for (int d = 0; d < 2; ++d) { t.mins[d] = KDT_INF; t.maxes[d] = -1; }
	for (int i = 0; i < t.npts; ++i)
		for (int d = 0; d < 2; ++d) {
			const int c = kdt_key(t.idx[i], d);
			if (c < t.mins[d]) t.mins[d] = c;
			if (c > t.maxes[d]) t.maxes[d] = c;
		}
}

// Serial build (host tests).  `stack` is scratch of at least 3 * 64 ints; node 0 is the root.
KDT_HD void kdt_build(KdtTree& t, int* stack, int max_nodes)
{
	uint32_t* idx = t.idx;
	t.nnodes = 0; t.overflow = 0;
	kdt_root_box(t);
	if (t.npts == 0) return;
	int sp = 0;
	// entry: (start, end, (parent << 1) | side), -2 for the root
	stack[0] = 0; stack[1] = t.npts; stack[2] = -2; sp = 1;
	while (sp > 0) {
		--sp;
		const int s = stack[3 * sp], e = stack[3 * sp + 1], link = stack[3 * sp + 2];
		if (t.nnodes >= max_nodes) { t.overflow = 1; return; }
		const int me = t.nnodes++;
		if (link >= 0) { if (link & 1) t.nodes[link >> 1].b = (uint16_t)me; else t.nodes[link >> 1].a = (uint16_t)me; }
		KdtNode nd;
		nd.a = (uint16_t)s; nd.b = (uint16_t)e; nd.split = 0; nd.dim = -1;
		int p, dim, split;
		if (kdt_split(idx, s, e, p, dim, split)) {
			nd.dim = (int16_t)dim; nd.split = (int16_t)split;
			if (sp + 2 > 64) { t.overflow = 1; return; }
			stack[3 * sp] = p; stack[3 * sp + 1] = e; stack[3 * sp + 2] = (me << 1) | 1; ++sp;
			stack[3 * sp] = s; stack[3 * sp + 1] = p; stack[3 * sp + 2] = (me << 1); ++sp;
		}
		t.nodes[me] = nd;
	}
}

#ifdef __CUDACC__
// Level-parallel build by one CTA: the nodes of a level are independent, so thread i splits frontier entry i; the
// arrangement inside every node is the serial one.  `frontier` is scratch of 2 * cap entries of 3 uint16, cap >=
// 2 * (npts / (KDT_LEAFSIZE + 1)) + 2.  t.idx / t.npts / t.nx are set by the caller (idx filled), all threads call.
struct KdtFrontier { uint16_t s, e, node; };
__device__ __forceinline__ void kdt_build_cta(KdtTree& t, KdtFrontier* frontier, int cap, int max_nodes, int* s_cnt /* [2] shared */)
{
	const int tid = threadIdx.x, nt = blockDim.x;
	if (tid == 0) {
		t.overflow = 0; t.nnodes = t.npts > 0 ? 1 : 0;
		kdt_root_box(t);
		frontier[0].s = 0; frontier[0].e = (uint16_t)t.npts; frontier[0].node = 0;
		s_cnt[0] = t.npts > 0 ? 1 : 0; s_cnt[1] = 0;
	}
	__syncthreads();
	int cur = 0;
	for (int level = 0; level < 4096; ++level) {
		const int nfr = s_cnt[cur];
		if (nfr == 0) break;
		KdtFrontier* fc = frontier + cur * cap;
		KdtFrontier* fn = frontier + (cur ^ 1) * cap;
		for (int i = tid; i < nfr; i += nt) {
			const int s = fc[i].s, e = fc[i].e, me = fc[i].node;
			KdtNode nd;
			nd.a = (uint16_t)s; nd.b = (uint16_t)e; nd.split = 0; nd.dim = -1;
			int p, dim, split;
			if (kdt_split(t.idx, s, e, p, dim, split)) {
				const int a = atomicAdd(&t.nnodes, 2);
				const int j = atomicAdd(&s_cnt[cur ^ 1], 2);
				if (a + 2 > max_nodes || j + 2 > cap) t.overflow = 1;
				else {
					nd.a = (uint16_t)a; nd.b = (uint16_t)(a + 1); nd.dim = (int16_t)dim; nd.split = (int16_t)split;
					fn[j].s = (uint16_t)s; fn[j].e = (uint16_t)p; fn[j].node = (uint16_t)a;
					fn[j + 1].s = (uint16_t)p; fn[j + 1].e = (uint16_t)e; fn[j + 1].node = (uint16_t)(a + 1);
				}
			}
			t.nodes[me] = nd;
		}
		__syncthreads();
		if (tid == 0) { s_cnt[cur] = 0; if (t.overflow) s_cnt[cur ^ 1] = 0; }
		cur ^= 1;
		__syncthreads();
	}
}
#endif

// ---- query --------------------------------------------------------------------------------------------------
struct KdtCell { int dist, sd0, sd1, node; };

// binary min-heap with scipy's sift rules (the tie behaviour is part of the traversal order)
template <typename Item>
KDT_HD void kdt_heap_push(Item* h, int& n, const Item& it)
{
	int i = n++;
	h[i] = it;
	while (i > 0 && h[i].dist < h[(i - 1) / 2].dist) {
		const Item t = h[(i - 1) / 2]; h[(i - 1) / 2] = h[i]; h[i] = t;
		i = (i - 1) / 2;
	}
}
template <typename Item>
KDT_HD void kdt_heap_remove(Item* h, int& n)
{
	h[0] = h[n - 1];
	--n;
	int i = 0, j = 1, k = 2;
	while ((j < n && h[i].dist > h[j].dist) || (k < n && h[i].dist > h[k].dist)) {
		const int l = (k < n && h[j].dist > h[k].dist) ? k : j;
		const Item t = h[l]; h[l] = h[i]; h[i] = t;
		i = l; j = 2 * i + 1; k = 2 * i + 2;
	}
}

struct KdtNb { int dist; int id; };   // dist = -(squared distance): the heap root is the current worst neighbour

// The k nearest good meshes of position (qy, qx) in ascending order of distance (the order the reference's query
// returns them in).  Returns their number (min(k, npts)); sets *overflow when the cell queue is exhausted.
KDT_HD int kdt_query(const KdtTree& t, int qy, int qx, int kmax, int* out_id, int* out_d2, int* overflow)
{
	KdtCell q[KDT_QCAP];
	KdtNb nb[KDT_K];
	int nq = 0, nn = 0;
	if (t.npts == 0) return 0;
	const int x[2] = {qy, qx};
	KdtCell cur;
	{
		int sd[2];
		for (int d = 0; d < 2; ++d) {
			int s = 0, u = x[d] - t.maxes[d];
			if (u > s) s = u; else { u = t.mins[d] - x[d]; if (u > s) s = u; }
			sd[d] = s * s;
		}
		cur.sd0 = sd[0]; cur.sd1 = sd[1]; cur.dist = sd[0] + sd[1]; cur.node = 0;
	}
	int upper = KDT_INF;
	for (;;) {
		const KdtNode nd = t.nodes[cur.node];
		if (nd.dim < 0) {
			for (int i = nd.a; i < nd.b; ++i) {
				const uint32_t w = t.idx[i];
				const int g = kdt_id(w);
				const int dy = kdt_key(w, 0) - qy, dx = kdt_key(w, 1) - qx;
				const int d2 = dy * dy + dx * dx;
				if (d2 < upper) {
					if (nn == kmax) kdt_heap_remove(nb, nn);
					KdtNb it; it.dist = -d2; it.id = g;
					kdt_heap_push(nb, nn, it);
					if (nn == kmax) upper = -nb[0].dist;
				}
			}
			if (nq == 0) break;
			cur = q[0];
			kdt_heap_remove(q, nq);
		} else {
			if (cur.dist > upper) break;
			KdtCell far = cur;
			int side;
			if (x[nd.dim] < nd.split) { cur.node = nd.a; far.node = nd.b; side = nd.split - x[nd.dim]; }
			else { cur.node = nd.b; far.node = nd.a; side = x[nd.dim] - nd.split; }
			side *= side;
			if (nd.dim == 0) { far.dist += side - far.sd0; far.sd0 = side; }
			else { far.dist += side - far.sd1; far.sd1 = side; }
			if (cur.dist > far.dist) { const KdtCell tmp = cur; cur = far; far = tmp; }
			if (far.dist <= upper) {
				if (nq >= KDT_QCAP) { *overflow = 1; }
				else kdt_heap_push(q, nq, far);
			}
		}
	}
	const int m = nn;
	for (int i = m - 1; i >= 0; --i) {
		out_id[i] = nb[0].id; out_d2[i] = -nb[0].dist;
		kdt_heap_remove(nb, nn);
	}
	return m;
}

// Shepard interpolation of one position from its neighbours (photutils ShepardIDWInterpolator.__call__ with power = 1,
// reg = 0, conf_dist = 1e-12): a coincident point returns its value, otherwise sum(w v) / sum(w) with w = 1 / d, both sums
// in the order numpy's pairwise np.sum takes for <= 10 terms (eight strided accumulators, then the rest in sequence).
template <typename ValueOf>
KDT_HD double kdt_shepard(const int* id, const int* d2, int m, ValueOf value_of)
{
	if (m == 0) return (double)NAN;
	for (int j = 0; j < m; ++j) if (d2[j] == 0) return value_of(id[j]);
	double w[KDT_K], wv[KDT_K];
	for (int j = 0; j < m; ++j) { w[j] = 1.0 / sqrt((double)d2[j]); wv[j] = w[j] * value_of(id[j]); }
	double sw, swv;
	if (m < 8) {
		sw = 0.0; swv = 0.0;
		for (int j = 0; j < m; ++j) { sw += w[j]; swv += wv[j]; }
	} else {
		sw = ((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]));
		swv = ((wv[0] + wv[1]) + (wv[2] + wv[3])) + ((wv[4] + wv[5]) + (wv[6] + wv[7]));
		for (int j = 8; j < m; ++j) { sw += w[j]; swv += wv[j]; }
	}
	return swv / sw;
}
