# /* **************************************************************************
#  *                                                                          *
#  *     (C) Copyright Paul Mensonides 2002-2011.                             *
#  *     (C) Copyright Edward Diener 2011.                                    *
#  *     Distributed under the Boost Software License, Version 1.0. (See      *
#  *     accompanying file LICENSE_1_0.txt or copy at                         *
#  *     http://www.boost.org/LICENSE_1_0.txt)                                *
#  *                                                                          *
#  ************************************************************************** */
#
# /* See http://www.boost.org for most recent version. */
#
# ifndef BOOST_PREPROCESSOR_SEQ_HPP
# define BOOST_PREPROCESSOR_SEQ_HPP
#
# include <libint2/boost/preprocessor/seq/cat.hpp>
# include <libint2/boost/preprocessor/seq/elem.hpp>
# include <libint2/boost/preprocessor/seq/enum.hpp>
# include <libint2/boost/preprocessor/seq/filter.hpp>
# include <libint2/boost/preprocessor/seq/first_n.hpp>
# include <libint2/boost/preprocessor/seq/fold_left.hpp>
# include <libint2/boost/preprocessor/seq/fold_right.hpp>
# include <libint2/boost/preprocessor/seq/for_each.hpp>
# include <libint2/boost/preprocessor/seq/for_each_i.hpp>
# include <libint2/boost/preprocessor/seq/for_each_product.hpp>
# include <libint2/boost/preprocessor/seq/insert.hpp>
# include <libint2/boost/preprocessor/seq/pop_back.hpp>
# include <libint2/boost/preprocessor/seq/pop_front.hpp>
# include <libint2/boost/preprocessor/seq/push_back.hpp>
# include <libint2/boost/preprocessor/seq/push_front.hpp>
# include <libint2/boost/preprocessor/seq/remove.hpp>
# include <libint2/boost/preprocessor/seq/replace.hpp>
# include <libint2/boost/preprocessor/seq/rest_n.hpp>
# include <libint2/boost/preprocessor/seq/reverse.hpp>
# include <libint2/boost/preprocessor/seq/seq.hpp>
# include <libint2/boost/preprocessor/seq/size.hpp>
# include <libint2/boost/preprocessor/seq/subseq.hpp>
# include <libint2/boost/preprocessor/seq/to_array.hpp>
# include <libint2/boost/preprocessor/seq/to_list.hpp>
# include <libint2/boost/preprocessor/seq/to_tuple.hpp>
# include <libint2/boost/preprocessor/seq/transform.hpp>
# include <libint2/boost/preprocessor/seq/variadic_seq_to_seq.hpp>
#
# endif
