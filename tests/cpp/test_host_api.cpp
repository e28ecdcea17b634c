// C++ host-API tests over libtncb200, written to read like the reference's own tests
// (tnc/src/tensornetwork/contraction.rs:226-264, io/qasm/qasm_importer.rs:171-194,
// builders/circuit_builder.rs:372-396).  Needs a GPU; run by tests/test_gpu_cpp_host.py.
#include <cmath>
#include <cstdio>
#include "tnc.hpp"

using namespace tnc;
static int failures = 0;
#define EXPECT(cond) do { if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)
static bool approx(Complex64 a, Complex64 b, double eps = 4 * 2.220446049250313e-16) { return std::abs(a - b) <= eps; }

static void test_outer_product_contraction(Context& ctx) {
  Tensor t1({0}, {3}), t2({1}, {2});
  t1.set_tensor_data(TensorData::new_from_data({3}, {{1, 0}, {2, 5}, {3, -1}}));
  t2.set_tensor_data(TensorData::new_from_data({2}, {{-4, 2}, {0, -1}}));
  Tensor t3 = Tensor::new_composite({t1, t2});
  Tensor result = contract_tensor_network(ctx, std::move(t3), ContractionPath::single(0, 1));
  EXPECT((result.legs == std::vector<uint64_t>{1, 0}));
  EXPECT((result.bond_dims == std::vector<uint64_t>{2, 3}));
  const Complex64 ref[6] = {{-4, 2}, {-18, -16}, {-10, 10}, {0, -1}, {5, -2}, {-1, -3}};
  auto e = result.elements();
  for (int i = 0; i < 6; i++) EXPECT(e[i] == ref[i]);
}

static Tensor ket0(uint64_t edge) { Tensor t = Tensor::new_from_const({edge}, 2); t.set_tensor_data(TensorData::new_from_data({2}, {{1, 0}, {0, 0}})); return t; }
static Tensor gate(std::vector<uint64_t> legs, const char* name) { Tensor t = Tensor::new_from_const(std::move(legs), 2); t.set_tensor_data(TensorData::gate(name)); return t; }

static void test_bell_contract(Context& ctx) {
  // qreg q[2]; h q[0]; cx q[0], q[1];  statevector network, legs as Circuit::append_gate numbers them
  Tensor tn = Tensor::new_composite({ket0(0), ket0(1), gate({0, 2}, "h"), gate({2, 1, 3, 4}, "cx")});
  Tensor r = contract_tensor_network(ctx, std::move(tn), ContractionPath::simple({{0, 1}, {0, 2}, {0, 3}}));
  EXPECT((r.legs == std::vector<uint64_t>{3, 4}));
  auto e = r.elements();
  const double h = 0.70710678118654752440;
  EXPECT(approx(e[0], {h, 0}) && approx(e[1], {0, 0}) && approx(e[2], {0, 0}) && approx(e[3], {h, 0}));
}

static void test_hadamards_amplitude(Context& ctx) {
  std::vector<Tensor> ts;
  const int qubits = 5;
  for (int q = 0; q < qubits; q++) ts.push_back(ket0(q));
  for (int q = 0; q < qubits; q++) ts.push_back(gate({(uint64_t)q, (uint64_t)(qubits + q)}, "h"));
  for (int q = 0; q < qubits; q++) ts.push_back(ket0(qubits + q));  // <0| bras
  std::vector<std::pair<size_t, size_t>> p;
  for (size_t i = 1; i < ts.size(); i++) p.push_back({0, i});
  Tensor r = contract_tensor_network(ctx, Tensor::new_composite(ts), ContractionPath::simple(p));
  EXPECT(r.legs.empty());
  EXPECT(approx(r.elements()[0], {std::pow(0.70710678118654752440, qubits), 0}));
}

static void test_panics_become_errors(Context& ctx) {
  Tensor tn = Tensor::new_composite({ket0(0), gate({0, 1}, "h"), ket0(1)});
  try { contract_tensor_network(ctx, tn, ContractionPath::simple({{0, 1}, {2, 1}})); EXPECT(false); }
  catch (const Error& e) { EXPECT(e.status == TNCB_ERR_UNCONTRACTED); }
  try { contract_tensor_network(ctx, tn, ContractionPath::simple({{0, 1}})); EXPECT(false); }
  catch (const Error& e) { EXPECT(e.status == TNCB_ERR_NOT_CONTRACTED); }
  try { contract_tensor_network(ctx, Tensor::new_composite({ket0(0), gate({0, 1}, "foo"), ket0(1)}), ContractionPath::simple({{0, 1}, {0, 2}})); EXPECT(false); }
  catch (const Error& e) { EXPECT(e.status == TNCB_ERR_GATE && std::string(e.what()) == "Gate 'foo' not found."); }
}

static void test_nested(Context& ctx) {
  Tensor a = Tensor::new_composite({ket0(0), gate({0, 1}, "h")});
  Tensor b = Tensor::new_composite({ket0(1)});
  ContractionPath p; p.nested[0] = ContractionPath::single(0, 1); p.nested[1] = ContractionPath(); p.toplevel = {{0, 1}};
  Tensor r = contract_tensor_network(ctx, Tensor::new_composite({a, b}), p);
  EXPECT(approx(r.elements()[0], {0.70710678118654752440, 0}));
}

static void test_plan_and_repeated_calls(Context& ctx) {
  // the same 5-qubit Hadamard amplitude through a NetworkPlan (execute, stage + run) and through repeated direct calls
  // (second sighting compiles a cached plan): every route gives the identical value
  std::vector<Tensor> ts;
  const int qubits = 5;
  for (int q = 0; q < qubits; q++) ts.push_back(ket0(q));
  for (int q = 0; q < qubits; q++) ts.push_back(gate({(uint64_t)q, (uint64_t)(qubits + q)}, "h"));
  for (int q = 0; q < qubits; q++) ts.push_back(ket0(qubits + q));
  std::vector<std::pair<size_t, size_t>> p;
  for (size_t i = 1; i < ts.size(); i++) p.push_back({0, i});
  const Tensor tn = Tensor::new_composite(ts);
  const ContractionPath path = ContractionPath::simple(p);
  const Complex64 direct = contract_tensor_network(ctx, tn, path).elements()[0];
  NetworkPlan plan(ctx, tn, path);
  EXPECT(plan.execute(tn).elements()[0] == direct);
  plan.stage(tn);
  EXPECT(plan.run().elements()[0] == direct);
  EXPECT(plan.run().elements()[0] == direct);
  for (int rep = 0; rep < 3; rep++) EXPECT(contract_tensor_network(ctx, tn, path).elements()[0] == direct);
  ctx.set_tolerance(1e-10); ctx.set_tolerance(0.0);
}

int main() {
  try {
    Context ctx(0);
    test_outer_product_contraction(ctx);
    test_bell_contract(ctx);
    test_hadamards_amplitude(ctx);
    test_panics_become_errors(ctx);
    test_nested(ctx);
    test_plan_and_repeated_calls(ctx);
  } catch (const Error& e) { std::printf("FAIL uncaught tnc::Error %d: %s\n", e.status, e.what()); return 2; }
  if (failures) { std::printf("%d failure(s)\n", failures); return 1; }
  std::printf("HOST_API_OK 6 tests\n");
  return 0;
}
