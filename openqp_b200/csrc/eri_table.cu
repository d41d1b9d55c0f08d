#include "eri_kernel.cuh"
namespace oqpb {
void fill_class_table_0(ClassEntry* t);
void fill_class_table_1(ClassEntry* t);
void fill_class_table_2(ClassEntry* t);
void fill_class_table_3(ClassEntry* t);
void fill_class_table_4(ClassEntry* t);
void fill_class_table_5(ClassEntry* t);
void fill_class_table_6(ClassEntry* t);
void fill_class_table_7(ClassEntry* t);
void fill_class_table_8(ClassEntry* t);
void fill_class_table_9(ClassEntry* t);
const ClassEntry* class_table() {
  static ClassEntry tab[55];
  static bool init = false;
  if (!init) {
    fill_class_table_0(tab);
    fill_class_table_1(tab);
    fill_class_table_2(tab);
    fill_class_table_3(tab);
    fill_class_table_4(tab);
    fill_class_table_5(tab);
    fill_class_table_6(tab);
    fill_class_table_7(tab);
    fill_class_table_8(tab);
    fill_class_table_9(tab);
    init = true;
  }
  return tab;
}
}
