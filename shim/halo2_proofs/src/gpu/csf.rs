//! `plonk::ConstraintSystem<Fr>` -> CSF blob (include/zkb200.h: "ZSF1" header, flattened expression DAG, gates in constraint-system
//! order, chunked mv-lookups, permutation columns, query lists).  Selectors are already compiled into fixed columns by keygen and
//! `chunk_lookups()` has been applied by the circuit's `configure` (super_circuit/test.rs:59, aggregator/src/aggregation/config.rs:223).
use crate::plonk::{Any, ConstraintSystem, Expression};
use halo2curves::bn256::Fr;
use std::collections::HashMap;

const MAGIC: u32 = 0x3146_535a;
#[repr(u32)]
enum Op { Const = 0, Fixed = 1, Advice = 2, Instance = 3, Challenge = 4, Neg = 5, Add = 6, Mul = 7, Scaled = 8 }

#[derive(Default)]
struct Dag {
    nodes: Vec<[u32; 3]>,
    consts: Vec<Fr>,
    const_ix: HashMap<[u8; 32], u32>,
}

impl Dag {
    fn constant(&mut self, c: Fr) -> u32 {
        use ff::PrimeField;
        let key = c.to_repr();
        if let Some(&i) = self.const_ix.get(&key) {
            return i;
        }
        self.consts.push(c);
        self.const_ix.insert(key, (self.consts.len() - 1) as u32);
        (self.consts.len() - 1) as u32
    }
    fn push(&mut self, n: [u32; 3]) -> u32 {
        self.nodes.push(n);
        (self.nodes.len() - 1) as u32
    }
    /// children precede parents; rotations are stored as two's-complement i32
    fn visit(&mut self, e: &Expression<Fr>) -> u32 {
        match e {
            Expression::Constant(c) => { let i = self.constant(*c); self.push([Op::Const as u32, i, 0]) }
            Expression::Selector(_) => unreachable!("selectors are compiled into fixed columns before proving"),
            Expression::Fixed(q) => self.push([Op::Fixed as u32, q.column_index() as u32, q.rotation().0 as u32]),
            Expression::Advice(q) => self.push([Op::Advice as u32, q.column_index() as u32, q.rotation().0 as u32]),
            Expression::Instance(q) => self.push([Op::Instance as u32, q.column_index() as u32, q.rotation().0 as u32]),
            Expression::Challenge(c) => self.push([Op::Challenge as u32, c.index() as u32, 0]),
            Expression::Negated(a) => { let a = self.visit(a); self.push([Op::Neg as u32, a, 0]) }
            Expression::Sum(a, b) => { let (a, b) = (self.visit(a), self.visit(b)); self.push([Op::Add as u32, a, b]) }
            Expression::Product(a, b) => { let (a, b) = (self.visit(a), self.visit(b)); self.push([Op::Mul as u32, a, b]) }
            Expression::Scaled(a, c) => { let a = self.visit(a); let i = self.constant(*c); self.push([Op::Scaled as u32, a, i]) }
        }
    }
}

pub fn encode(cs: &ConstraintSystem<Fr>, k: u32) -> Vec<u32> {
    let mut dag = Dag::default();
    let gates: Vec<u32> = cs.gates().iter().flat_map(|g| g.polynomials().iter()).map(|p| dag.visit(p)).collect();
    let lookups: Vec<(Vec<Vec<u32>>, Vec<u32>)> = cs
        .lookups()
        .iter()
        .map(|l| {
            let inputs = l.inputs_expressions().iter().map(|set| set.iter().map(|e| dag.visit(e)).collect()).collect();
            let table = l.table_expressions().iter().map(|e| dag.visit(e)).collect();
            (inputs, table)
        })
        .collect();
    let perm = cs.permutation().get_columns();
    let mut w: Vec<u32> = vec![
        MAGIC, k, cs.num_fixed_columns() as u32, cs.num_advice_columns() as u32, cs.num_instance_columns() as u32, cs.num_challenges() as u32,
        cs.blinding_factors() as u32, cs.degree() as u32, cs.phases().count() as u32, dag.nodes.len() as u32, dag.consts.len() as u32,
        gates.len() as u32, lookups.len() as u32, perm.len() as u32, cs.advice_queries().len() as u32, cs.fixed_queries().len() as u32,
        cs.instance_queries().len() as u32, 0,
    ];
    w.extend(cs.advice_column_phase().iter().map(|&p| p as u32));
    w.extend(cs.challenge_phase().iter().map(|&p| p as u32));
    for n in &dag.nodes { w.extend_from_slice(n); }
    for c in &dag.consts {
        let limbs: [u64; 4] = unsafe { std::mem::transmute_copy(c) };   // Montgomery limbs, as in memory
        for l in limbs { w.push(l as u32); w.push((l >> 32) as u32); }
    }
    w.extend(&gates);
    for (inputs, table) in &lookups {
        w.push(inputs.len() as u32);
        w.push(table.len() as u32);
        for set in inputs { w.extend(set); }
        w.extend(table);
    }
    for c in perm {
        let ty = match c.column_type() { Any::Fixed => 1, Any::Advice(_) => 2, Any::Instance => 3 };
        w.push(ty);
        w.push(c.index() as u32);
    }
    for (c, r) in cs.advice_queries() { w.push(c.index() as u32); w.push(r.0 as u32); }
    for (c, r) in cs.fixed_queries() { w.push(c.index() as u32); w.push(r.0 as u32); }
    for (c, r) in cs.instance_queries() { w.push(c.index() as u32); w.push(r.0 as u32); }
    debug_assert_eq!(unsafe { crate::zkb200_sys::zkb_csf_validate(w.as_ptr(), w.len() as u64) }, 0);
    w
}
